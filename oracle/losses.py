"""Oracle: loss layer (reference ``loss_functions.py``).  TEST INFRASTRUCTURE.

fp32 torch restatement; gradients come from autograd of this restatement and are
pinned against the reference's own autograd by tests/golden (see oracle/__init__)."""
import torch
import torch.nn.functional as F
from .geometry import inverse_warp, flow_warp, pose2flow
from .ssim import ssim

epsilon = 1e-8


def _pool(img, h, w):
    return F.adaptive_avg_pool2d(img, (h, w))


def spatial_normalize(disp):
    """Reference loss_functions.py:13-16."""
    m = disp.mean(dim=1, keepdim=True).mean(dim=2, keepdim=True).mean(dim=3, keepdim=True)
    return disp / m


def robust_l1_per_pix(x, q=0.5, eps=1e-2):
    """(x^2+eps)^q.  Reference loss_functions.py:23-25."""
    return torch.pow(x.pow(2) + eps, q)


def robust_l1(x, q=0.5, eps=1e-2):
    """Reference loss_functions.py:18-21."""
    return robust_l1_per_pix(x, q, eps).mean()


def occlusion_masks(flow_bw, flow_fw):
    """Two identical masks.  Reference loss_functions.py:343-352."""
    mag_sq = flow_fw.pow(2).sum(dim=1) + flow_bw.pow(2).sum(dim=1)
    s = (flow_fw + flow_bw).sum(dim=1)
    occ = (s > 0.08 * mag_sq + 1.0).type_as(flow_bw)
    return occ, occ


def depth_occlusion_masks(depth, pose, intrinsics, intrinsics_inv):
    """Uses the UNSCALED intrinsics at every level.
    Reference loss_functions.py:132-137."""
    fc = [pose2flow(depth.squeeze(), pose[:, i], intrinsics, intrinsics_inv)
          for i in range(pose.size(1))]
    m1, m2 = occlusion_masks(fc[1], fc[2])
    m0, m3 = occlusion_masks(fc[0], fc[3])
    return torch.stack((m0, m1, m2, m3), dim=1)


def photometric_reconstruction_loss(tgt_img, ref_imgs, intrinsics, intrinsics_inv, depth,
                                    explainability_mask, pose, rotation_mode='euler',
                                    padding_mode='zeros', lambda_oob=0, qch=0.5, wssim=0.5):
    """Reference loss_functions.py:80-128."""
    def one_scale(depth, mask, occ):
        assert mask is None or depth.size()[2:] == mask.size()[2:]
        assert pose.size(1) == len(ref_imgs)
        loss = 0
        b, _, h, w = depth.size()
        downscale = tgt_img.size(2) / h
        tgt_s = _pool(tgt_img, h, w)
        refs_s = [_pool(r, h, w) for r in ref_imgs]
        K_s = torch.cat((intrinsics[:, 0:2] / downscale, intrinsics[:, 2:]), dim=1)
        Kinv_s = torch.cat((intrinsics_inv[:, :, 0:2] * downscale, intrinsics_inv[:, :, 2:]), dim=2)
        for i, ref in enumerate(refs_s):
            warped = inverse_warp(ref, depth[:, 0], pose[:, i], K_s, Kinv_s, rotation_mode, padding_mode)
            valid = 1 - (warped == 0).prod(1, keepdim=True).type_as(warped)
            diff = (tgt_s - warped) * valid
            ssim_loss = 1 - ssim(tgt_s, warped) * valid
            oob = valid.nelement() / valid.sum()
            if mask is not None:
                diff = diff * (1 - occ[:, i:i + 1]) * mask[:, i:i + 1].expand_as(diff)
                ssim_loss = ssim_loss * (1 - occ[:, i:i + 1]) * mask[:, i:i + 1].expand_as(ssim_loss)
            else:
                diff = diff * (1 - occ[:, i:i + 1]).expand_as(diff)
                ssim_loss = ssim_loss * (1 - occ[:, i:i + 1]).expand_as(ssim_loss)
            loss = loss + (1 - wssim) * oob * (robust_l1(diff, q=qch) + wssim * ssim_loss.mean()) \
                + lambda_oob * robust_l1(1 - valid, q=qch)
        return loss

    if type(explainability_mask) not in [tuple, list]:
        explainability_mask = [explainability_mask]
    if type(depth) not in [list, tuple]:
        depth = [depth]
    loss = 0
    for d, m in zip(depth, explainability_mask):
        occ = depth_occlusion_masks(d, pose, intrinsics, intrinsics_inv)
        loss = loss + one_scale(d, m, occ)
    return loss


def photometric_flow_loss(tgt_img, ref_imgs, flows, explainability_mask, lambda_oob=0, qch=0.5,
                          wssim=0.5):
    """Reference loss_functions.py:27-77."""
    def one_scale(mask, occ, flows):
        assert mask is None or flows[0].size()[2:] == mask.size()[2:]
        assert len(flows) == len(ref_imgs)
        loss = 0
        b, _, h, w = flows[0].size()
        tgt_s = _pool(tgt_img, h, w)
        refs_s = [_pool(r, h, w) for r in ref_imgs]
        for i, ref in enumerate(refs_s):
            warped = flow_warp(ref, flows[i])
            valid = 1 - (warped == 0).prod(1, keepdim=True).type_as(warped)
            diff = (tgt_s - warped) * valid
            ssim_loss = 1 - ssim(tgt_s, warped) * valid
            oob = valid.nelement() / valid.sum()
            if mask is not None:
                diff = diff * mask[:, i:i + 1].expand_as(diff)
                ssim_loss = ssim_loss * mask[:, i:i + 1].expand_as(ssim_loss)
            if occ is not None:
                diff = diff * (1 - occ[:, i:i + 1]).expand_as(diff)
                ssim_loss = ssim_loss * (1 - occ[:, i:i + 1]).expand_as(ssim_loss)
            loss = loss + (1 - wssim) * oob * (robust_l1(diff, q=qch) + wssim * ssim_loss.mean()) \
                + lambda_oob * robust_l1(1 - valid, q=qch)
        return loss

    if type(flows[0]) not in [tuple, list]:
        if explainability_mask is not None:
            explainability_mask = [explainability_mask]
        flows = [[uv] for uv in flows]
    loss = 0
    for i in range(len(flows[0])):
        fl = [uv[i] for uv in flows]
        occ_bw, occ_fw = occlusion_masks(fl[0], fl[1])
        occ = torch.stack((occ_bw, occ_fw), dim=1)
        loss = loss + one_scale(explainability_mask[i], occ, fl)
    return loss


def gaussian_explainability_loss(mask):
    """Reference loss_functions.py:139-145."""
    if type(mask) not in [tuple, list]:
        mask = [mask]
    loss = 0
    for m in mask:
        loss = loss + torch.exp(-torch.mean((m - 0.5).pow(2)) / 0.15)
    return loss


def explainability_loss(mask):
    """BCE(mask, 1) per level.  Reference loss_functions.py:148-155."""
    if type(mask) not in [tuple, list]:
        mask = [mask]
    loss = 0
    for m in mask:
        loss = loss + F.binary_cross_entropy(m, torch.ones_like(m))
    return loss


def logical_or(a, b):
    """Reference loss_functions.py:157-158."""
    return 1 - (1 - a) * (1 - b)


def consensus_sides(cam_flows_fwd, cam_flows_bwd, flows_fwd, flows_bwd, tgt_img, ref_img_fwd,
                    ref_img_bwd, wssim, wrig, ws=0.1):
    """The two sides of the consensus comparison per level: (wrig * cam_err, flow_err).
    Reference loss_functions.py:160-198 (the comparison itself is :199-200)."""
    def valid_of(w):
        return 1 - (w == 0).prod(1, keepdim=True).type_as(w)

    def err(tgt, w):
        return (1 - wssim) * robust_l1_per_pix(tgt - w).mean(1, keepdim=True) \
            + wssim * (1 - ssim(tgt, w)).mean(1, keepdim=True)

    out = []
    for i in range(len(cam_flows_fwd)):
        b, _, h, w = cam_flows_fwd[i].size()
        tgt = _pool(tgt_img, h, w)
        rf, rb = _pool(ref_img_fwd, h, w), _pool(ref_img_bwd, h, w)
        cam_f, cam_b = flow_warp(rf, cam_flows_fwd[i]), flow_warp(rb, cam_flows_bwd[i])
        flo_f = flow_warp(rf, flows_fwd[i])
        valid_cam = logical_or(valid_of(cam_f), valid_of(cam_b))
        cam_err = torch.min(err(tgt, cam_f), err(tgt, cam_b)) * valid_cam
        flow_err = err(tgt, flo_f)
        out.append((wrig * cam_err, flow_err))
    return out


def consensus_exp_masks(cam_flows_fwd, cam_flows_bwd, flows_fwd, flows_bwd, tgt_img, ref_img_fwd,
                        ref_img_bwd, wssim, wrig, ws=0.1):
    """0/1 targets, no grad.  Reference loss_functions.py:160-202."""
    sides = consensus_sides(cam_flows_fwd, cam_flows_bwd, flows_fwd, flows_bwd, tgt_img, ref_img_fwd,
                            ref_img_bwd, wssim, wrig, ws)
    return [(lhs <= (rhs + epsilon)).type_as(lhs) for lhs, rhs in sides]


def compute_joint_mask_for_depth(explainability_mask, rigidity_mask_bwd, rigidity_mask_fwd, THRESH):
    """Reference loss_functions.py:204-219."""
    joint = []
    for i in range(len(explainability_mask)):
        e = explainability_mask[i]
        rf = (rigidity_mask_fwd[i] > THRESH).type_as(e)
        rb = (rigidity_mask_bwd[i] > THRESH).type_as(e)
        ej = (1 - (1 - e[:, 1]) * (1 - e[:, 2]).unsqueeze(1) > 0.5).type_as(e)
        jf = logical_or(rf, ej).detach()
        jb = logical_or(rb, ej).detach()
        joint.append(torch.cat((jb, jb, jf, jf), dim=1))
    return joint


def weighted_binary_cross_entropy(output, target, weights=None):
    """Reference loss_functions.py:252-261."""
    if weights is not None:
        assert len(weights) == 2
        loss = weights[1] * (target * torch.log(output + epsilon)) + \
            weights[0] * ((1 - target) * torch.log(1 - output + epsilon))
    else:
        loss = target * torch.log(output + epsilon) + (1 - target) * torch.log(1 - output + epsilon)
    return torch.neg(torch.mean(loss))


def consensus_depth_flow_mask(explainability_mask, census_mask_bwd, census_mask_fwd,
                              exp_masks_bwd_target, exp_masks_fwd_target, THRESH, wbce):
    """Reference loss_functions.py:221-250."""
    assert len(explainability_mask) == len(census_mask_bwd)
    assert len(explainability_mask) == len(census_mask_fwd)
    loss = 0.
    for i in range(len(explainability_mask)):
        e = explainability_mask[i]
        cf = (census_mask_fwd[i] < THRESH).type_as(e).prod(dim=1, keepdim=True)
        cb = (census_mask_bwd[i] < THRESH).type_as(e).prod(dim=1, keepdim=True)
        cf = logical_or(cf, exp_masks_fwd_target[i]).detach()
        cb = logical_or(cb, exp_masks_bwd_target[i]).detach()
        tgt = torch.cat((cb, cb, cf, cf), dim=1)
        loss = loss + weighted_binary_cross_entropy(e, tgt.type_as(e), [wbce, 1 - wbce])
    return loss


def edge_aware_smoothness_loss(img, pred_disp):
    """Reference loss_functions.py:287-319 (per-level weight stays 1)."""
    def gx(t):
        return t[:, :, :-1, :] - t[:, :, 1:, :]

    def gy(t):
        return t[:, :, :, :-1] - t[:, :, :, 1:]

    loss = 0
    for p in pred_disp:
        b, _, h, w = p.size()
        im = _pool(img, h, w)
        wx = torch.exp(-torch.mean(torch.abs(gx(im)), 1, keepdim=True))
        wy = torch.exp(-torch.mean(torch.abs(gy(im)), 1, keepdim=True))
        loss = loss + torch.mean(torch.abs(gx(p)) * wx) + torch.mean(torch.abs(gy(p)) * wy)
    return loss


def smooth_loss(pred_disp):
    """2nd-order smoothness, weight /2.3 per level.  Reference loss_functions.py:323-341."""
    def grad(p):
        return p[:, :, :, 1:] - p[:, :, :, :-1], p[:, :, 1:] - p[:, :, :-1]

    if type(pred_disp) not in [tuple, list]:
        pred_disp = [pred_disp]
    loss = 0
    weight = 1.
    for p in pred_disp:
        dx, dy = grad(p)
        dx2, dxdy = grad(dx)
        dydx, dy2 = grad(dy)
        loss = loss + (dx2.abs().mean() + dxdy.abs().mean() + dydx.abs().mean() + dy2.abs().mean()) * weight
        weight /= 2.3
    return loss
