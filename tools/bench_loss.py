#!/usr/bin/env python
"""Developer micro-benchmark of the fused loss-layer kernels (not the driver's bench.py):
per-kernel CUDA-event timings at BASELINE.json's loss size (b4, 256x832, 6 levels), L2 flushed between
iterations, algorithmic GB/s per SURVEY.md 8(d)."""
import argparse
import json
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cc_b200 import synth, loss_functions as CL, pyramid   # noqa: E402


def time_fn(fn, iters, flush):
    evs = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--B', type=int, default=4)
    ap.add_argument('--iters', type=int, default=20)
    ap.add_argument('--levels', type=int, default=6)
    args = ap.parse_args()
    dev = torch.device('cuda:0')
    B, H, W, NL = args.B, 256, 832, args.levels
    s = synth.sample(B, H, W, seed=0, nlevels=NL)
    s = {k: ([t.to(dev) for t in v] if isinstance(v, list) else v.to(dev)) for k, v in s.items()}
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    px = sum(B * (H >> l) * (W >> l) for l in range(NL))
    res = {}
    for name, wssim, use_mask in (('rigid_ssim_mask', 0.997, True), ('rigid_ssim_nomask', 0.997, False),
                                  ('rigid_nossim_mask', 0.0, True)):
        depth = [d.clone().requires_grad_(True) for d in s['depth']]
        pose = s['pose'].clone().requires_grad_(True)
        em = [m.clone().requires_grad_(True) for m in s['emask']] if use_mask else [None] * NL
        holder = {}

        def fwd():
            holder['l'] = CL.photometric_reconstruction_loss(s['tgt'], s['refs'], s['K'], s['Kinv'], depth, em, pose,
                                                             wssim=wssim)

        def bwd():
            torch.autograd.grad(holder['l'], depth + [pose] + (em if use_mask else []), retain_graph=True)

        for _ in range(3):
            fwd(); bwd()
        tf, tf_min = time_fn(fwd, args.iters, flush)
        fwd()
        tb, tb_min = time_fn(bwd, args.iters, flush)
        byt_f = px * (80 if use_mask else 64)
        byt_fb = px * (100 if use_mask else 68)
        res[name] = dict(fwd_ms=tf, bwd_ms=tb, fwd_min_ms=tf_min, bwd_min_ms=tb_min,
                         alg_GBps_fwd_bwd=byt_fb / ((tf + tb) * 1e-3) / 1e9, alg_GBps_fwd=byt_f / (tf * 1e-3) / 1e9)
    # flow loss + smoothness + consensus
    ff = [f.clone().requires_grad_(True) for f in s['flow_fwd']]
    fb = [f.clone().requires_grad_(True) for f in s['flow_bwd']]
    em = [m.clone().requires_grad_(True) for m in s['emask']]
    holder = {}

    def f_fwd():
        holder['l'] = CL.photometric_flow_loss(s['tgt'], s['refs'][1:3], [fb, ff], [1 - m[:, 1:3] for m in em], wssim=0.997)

    def f_bwd():
        torch.autograd.grad(holder['l'], ff + fb + em, retain_graph=True)
    for _ in range(3):
        f_fwd(); f_bwd()
    tf, _ = time_fn(f_fwd, args.iters, flush)
    f_fwd()
    tb, _ = time_fn(f_bwd, args.iters, flush)
    res['flow_ssim_mask'] = dict(fwd_ms=tf, bwd_ms=tb, alg_GBps_fwd_bwd=px * 84 / ((tf + tb) * 1e-3) / 1e9)

    def sm():
        l = CL.edge_aware_smoothness_loss(s['tgt'], depth) + CL.edge_aware_smoothness_loss(s['tgt'], ff) + \
            CL.edge_aware_smoothness_loss(s['tgt'], fb) + CL.edge_aware_smoothness_loss(s['tgt'], em)
        torch.autograd.grad(l, depth + ff + fb + em)
    for _ in range(3):
        sm()
    t, _ = time_fn(sm, args.iters, flush)
    res['edge_smooth_x4_fwd_bwd'] = dict(ms=t, alg_GBps=px * 84 / (t * 1e-3) / 1e9)

    def pyr():
        pyramid.clear()
        for im in [s['tgt']] + s['refs']:
            pyramid.get(im, NL)
    for _ in range(3):
        pyr()
    t, _ = time_fn(pyr, args.iters, flush)
    res['pyramid_5_frames'] = dict(ms=t, GBps=5 * B * 3 * H * W * 4 * (1 + 1 / 3) / (t * 1e-3) / 1e9)
    print(json.dumps(res, indent=1))


if __name__ == '__main__':
    main()
