"""Oracle: validation metrics (reference ``loss_functions.py:355-467``).  TEST INFRASTRUCTURE.
fp32 torch restatement, pinned against fixtures frozen from the reference (tests/golden/metrics_small.npz)."""
import torch
import torch.nn.functional as F

epsilon = 1e-8


def _up(gt, pred):
    """Reference :357-361: bilinear resize to the ground-truth size (align_corners False) and rescale to gt pixels."""
    hp, wp = pred.shape[2], pred.shape[3]
    hg, wg = gt.shape[2], gt.shape[3]
    up = F.interpolate(pred, size=(hg, wg), mode='bilinear', align_corners=False)
    return up[:, 0] * (wg / wp), up[:, 1] * (hg / hp)


def flow_diff(gt, pred):
    """Reference :355-365."""
    u, v = _up(gt, pred)
    return torch.sqrt(torch.pow(gt[:, 0] - u, 2) + torch.pow(gt[:, 1] - v, 2))


def compute_epe(gt, pred):
    """Reference :368-387."""
    epe = flow_diff(gt, pred)
    if gt.size(1) == 3:
        valid = gt[:, 2]
        return ((epe * valid).sum() / (valid.sum() + epsilon)).item()
    return (epe.sum() / (gt.size(0) * gt.size(2) * gt.size(3))).item()


def outlier_err(gt, pred, tau=(3, 0.05)):
    """Reference :389-407."""
    valid = gt[:, 2]
    epe = flow_diff(gt, pred) * valid
    mag = torch.sqrt(torch.pow(gt[:, 0], 2) + torch.pow(gt[:, 1], 2))
    e0 = (epe > tau[0]).type_as(epe)
    e1 = ((epe / (mag + epsilon)) > tau[1]).type_as(epe)
    return ((e0 * e1 * valid).sum() / (valid.sum() + epsilon)).item()


def compute_all_epes(gt, rigid_pred, non_rigid_pred, rigidity_mask, THRESH=0.5):
    """Reference :409-427."""
    m_pred = F.interpolate(rigidity_mask, size=rigid_pred.shape[2:], mode='bilinear', align_corners=False)
    m_gt = F.interpolate(rigidity_mask, size=gt.shape[2:], mode='bilinear', align_corners=False)
    non_rigid_pred = (m_pred <= THRESH).type_as(non_rigid_pred).expand_as(non_rigid_pred) * non_rigid_pred
    rigid_pred = (m_pred > THRESH).type_as(rigid_pred).expand_as(rigid_pred) * rigid_pred
    total = non_rigid_pred + rigid_pred
    gt_non_rigid = (m_gt <= THRESH).type_as(gt).expand_as(gt) * gt
    gt_rigid = (m_gt > THRESH).type_as(gt).expand_as(gt) * gt
    return [compute_epe(gt, total), compute_epe(gt_rigid, rigid_pred), compute_epe(gt_non_rigid, non_rigid_pred),
            outlier_err(gt, total)]


def compute_errors(gt, pred, crop=True):
    """Reference :430-467 (torch.median = lower median)."""
    B, H, W = gt.shape
    keep = torch.ones(H, W, dtype=torch.bool)
    if crop:
        keep = torch.zeros(H, W, dtype=torch.bool)
        keep[int(0.40810811 * H):int(0.99189189 * H), int(0.03594771 * W):int(0.96405229 * W)] = True
    sums = [0.0] * 6
    for g, p in zip(gt, pred):
        sel = (g > 0) & (g < 80) & keep
        g, p = g[sel], p[sel].clamp(1e-3, 80)
        p = p * torch.median(g) / torch.median(p)
        th = torch.max(g / p, p / g)
        vals = [torch.mean(torch.abs(g - p)), torch.mean(torch.abs(g - p) / g), torch.mean(((g - p) ** 2) / g),
                (th < 1.25).float().mean(), (th < 1.25 ** 2).float().mean(), (th < 1.25 ** 3).float().mean()]
        sums = [a + b for a, b in zip(sums, vals)]
    return [v / B for v in sums]
