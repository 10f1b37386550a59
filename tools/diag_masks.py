#!/usr/bin/env python
"""Diagnostic: full-size valid/occlusion mask agreement of the CUDA kernel vs the oracle run on CPU and
on the GPU (torch CUDA ops), and where the gradient outliers of the full-size loss test come from."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cc_b200 import synth, loss_functions as CL   # noqa: E402
from oracle import losses as OL, geometry as OG   # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device('cuda:0')
B, H, W, NL = 4, 256, 832, 6
s = synth.sample(B, H, W, seed=5, nlevels=NL)
sg = {k: ([t.to(dev) for t in v] if isinstance(v, list) else v.to(dev)) for k, v in s.items()}


def masks(ss, lvl):
    h, w = H >> lvl, W >> lvl
    ds = H / h
    K_s = torch.cat((ss['K'][:, 0:2] / ds, ss['K'][:, 2:]), dim=1)
    Kinv_s = torch.cat((ss['Kinv'][:, :, 0:2] * ds, ss['Kinv'][:, :, 2:]), dim=2)
    occ = OL.depth_occlusion_masks(ss['depth'][lvl], ss['pose'], ss['K'], ss['Kinv'])
    refs_s = [torch.nn.functional.adaptive_avg_pool2d(r, (h, w)) for r in ss['refs']]
    out = []
    for i in range(4):
        wimg = OG.inverse_warp(refs_s[i], ss['depth'][lvl][:, 0], ss['pose'][:, i], K_s, Kinv_s)
        valid = 1 - (wimg == 0).prod(1).type_as(wimg)
        out.append(valid * (1 - occ[:, i]))
    return torch.stack(out, 1)


depth = [d.clone().requires_grad_(True) for d in sg['depth']]
pose = sg['pose'].clone().requires_grad_(True)
l = CL.photometric_reconstruction_loss(sg['tgt'], sg['refs'], sg['K'], sg['Kinv'], depth, [None] * NL, pose, wssim=0.997)
vos = [t for t in l.grad_fn.keep if t.dim() == 4 and t.shape[1] == 4 and t.shape[0] == B]
for lvl in range(NL):
    h, w = H >> lvl, W >> lvl
    got = [t for t in vos if tuple(t.shape) == (B, 4, h, w)][0]
    mc = masks(s, lvl).to(dev)
    mg = masks(sg, lvl)
    print(f'level {lvl}: kernel vs CPU-oracle mismatches {(got != mc).sum().item()}, kernel vs GPU-oracle '
          f'{(got != mg).sum().item()}, CPU-oracle vs GPU-oracle {(mc != mg).sum().item()} of {got.numel()}')


def grads(mod, ss):
    dp = [d.clone().requires_grad_(True) for d in ss['depth']]
    po = ss['pose'].clone().requires_grad_(True)
    em = [m.clone().requires_grad_(True) for m in ss['emask']]
    lo = mod.photometric_reconstruction_loss(ss['tgt'], ss['refs'], ss['K'], ss['Kinv'], dp, em, po, wssim=0.997)
    return lo, torch.autograd.grad(lo, dp + [po] + em)


lk, gk = grads(CL, sg)
lc, gc = grads(OL, s)
lg, gg = grads(OL, sg)
print('loss kernel/cpu/gpu', lk.item(), lc.item(), lg.item())
names = [f'depth{i}' for i in range(NL)] + ['pose'] + [f'mask{i}' for i in range(NL)]
for n, a, b, c in zip(names, gk, gc, gg):
    b = b.to(dev)
    sc = b.abs().max().item()
    e1 = (a - b).abs() / sc
    e2 = (a - c).abs() / sc
    e3 = (b - c).abs() / sc
    print(f'{n:8s} scale {sc:.3e}  kernel-vs-cpu max {e1.max().item():.2e} n>1e-4: {(e1 > 1e-4).sum().item():6d} | '
          f'kernel-vs-gpuoracle max {e2.max().item():.2e} n>1e-4: {(e2 > 1e-4).sum().item():6d} | cpu-vs-gpu oracle max '
          f'{e3.max().item():.2e} n>1e-4 {(e3 > 1e-4).sum().item():6d}  of {a.numel()}')
