"""Deterministic synthetic inputs of the reference's sample contract
(``datasets/sequence_folders.py:51-61``: tgt_img, ref_imgs[4], K, K^-1), SURVEY.md 8(d).

Host-side (CPU torch) generator shared by the tests, ``bench.py`` and the golden-fixture
script, so every leg sees identical tensors."""
import math
import torch


def _smooth_field(gen, b, c, h, w, nwaves=6):
    ys = torch.arange(h, dtype=torch.float32).view(1, 1, h, 1) / max(h, 1)
    xs = torch.arange(w, dtype=torch.float32).view(1, 1, 1, w) / max(w, 1)
    out = torch.zeros(b, c, h, w)
    for _ in range(nwaves):
        fx = torch.rand(b, c, 1, 1, generator=gen) * 6 + 0.5
        fy = torch.rand(b, c, 1, 1, generator=gen) * 4 + 0.5
        ph = torch.rand(b, c, 1, 1, generator=gen) * 2 * math.pi
        amp = torch.rand(b, c, 1, 1, generator=gen) * 0.5 + 0.1
        out = out + amp * torch.sin(2 * math.pi * (fx * xs + fy * ys) + ph)
    return out / nwaves * 2.5


def frames(B, H, W, seed=0, n_refs=4, noise=0.05):
    """tgt [B,3,H,W] in [-1,1] + n_refs shifted/perturbed copies (t-2,t-1,t+1,t+2)."""
    g = torch.Generator().manual_seed(seed)
    base = _smooth_field(g, B, 3, H + 16, W + 32).clamp(-0.95, 0.95)
    shifts = [(-2, -6), (-1, -3), (1, 3), (2, 6)][:n_refs] if n_refs == 4 else [(-1, -3), (1, 3)][:n_refs]

    def crop(dy, dx):
        im = base[:, :, 8 + dy:8 + dy + H, 16 + dx:16 + dx + W]
        return (im + noise * (torch.rand(im.shape, generator=g) * 2 - 1)).clamp(-1, 1).contiguous()

    tgt = crop(0, 0)
    refs = [crop(dy, dx) for dy, dx in shifts]
    return tgt, refs


def intrinsics(B, H, W):
    """KITTI-like K scaled to (H, W): fx=721.54*W/1242 ...."""
    fx, fy = 721.54 * W / 1242.0, 721.54 * H / 375.0
    cx, cy = 609.56 * W / 1242.0, 172.85 * H / 375.0
    K = torch.tensor([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], dtype=torch.float32)
    K = K.unsqueeze(0).repeat(B, 1, 1)
    return K, torch.inverse(K)


def pyramid_sizes(H, W, nlevels=6):
    return [(H >> l, W >> l) for l in range(nlevels)]


def depths(B, H, W, nlevels=6, seed=1):
    """depth_l = 1/disp_l, disp = 10*sigmoid(z)+0.01 with smooth z; list of [B,1,h,w]."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for h, w in pyramid_sizes(H, W, nlevels):
        z = _smooth_field(g, B, 1, h, w) * 2 - 1.5
        out.append(1.0 / (10 * torch.sigmoid(z) + 0.01))
    return out


def poses(B, n_refs=4, seed=2, big_tx_sample=True):
    """[B,n_refs,6] ~ N(0,1)*0.01*[5,1,10,1,1,1]; sample 0 gets a large tx so >=5% px go OOB."""
    g = torch.Generator().manual_seed(seed)
    scale = torch.tensor([5, 1, 10, 1, 1, 1], dtype=torch.float32) * 0.01
    p = torch.randn(B, n_refs, 6, generator=g) * scale
    if big_tx_sample:
        p[0, :, 0] += 0.15
    return p


def flows(B, H, W, nlevels=6, seed=3):
    """list over levels of [B,2,h,w] ~ smooth N(0,(3*2^-l)^2) px."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for l, (h, w) in enumerate(pyramid_sizes(H, W, nlevels)):
        out.append(_smooth_field(g, B, 2, h, w) * 3.0 * (0.5 ** l) * 2
                   + 0.2 * (0.5 ** l) * torch.randn(B, 2, h, w, generator=g))
    return out


def exp_masks(B, H, W, ch=4, nlevels=6, seed=4):
    g = torch.Generator().manual_seed(seed)
    return [torch.sigmoid(_smooth_field(g, B, ch, h, w) * 2 + 0.5 * torch.randn(B, ch, h, w, generator=g))
            for h, w in pyramid_sizes(H, W, nlevels)]


def sample(B, H, W, seed=0, nlevels=6):
    """Everything one loss-layer call needs, CPU fp32."""
    tgt, refs = frames(B, H, W, seed=seed)
    K, Kinv = intrinsics(B, H, W)
    return dict(tgt=tgt, refs=refs, K=K, Kinv=Kinv,
                depth=depths(B, H, W, nlevels, seed + 1), pose=poses(B, 4, seed + 2),
                flow_fwd=flows(B, H, W, nlevels, seed + 3), flow_bwd=flows(B, H, W, nlevels, seed + 4),
                emask=exp_masks(B, H, W, 4, nlevels, seed + 5))


def seeded_fill(module, seed):
    """Deterministic parameters for ANY module from its state_dict key names and shapes alone - lets a fixture frozen
    from a reference module and a test on the mirrored module agree on the weights without storing them: conv weights
    ~ N(0, 1/fan), biases ~ 0.1 N(0,1), BatchNorm weight 1 + 0.1 N(0,1); running statistics keep their defaults."""
    import torch
    sd = module.state_dict()
    with torch.no_grad():
        for i, k in enumerate(sorted(sd)):
            t = sd[k]
            if not t.dtype.is_floating_point or k.endswith('running_mean') or k.endswith('running_var'):
                continue
            g = torch.Generator().manual_seed(seed * 100003 + i)
            if t.dim() == 4:
                v = torch.randn(t.shape, generator=g) / (t[0].numel() ** 0.5)
            elif k.endswith('weight'):                       # BatchNorm scale
                v = 1.0 + 0.1 * torch.randn(t.shape, generator=g)
            else:
                v = 0.1 * torch.randn(t.shape, generator=g)
            t.copy_(v.to(t.device))
    return module
