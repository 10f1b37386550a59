#!/usr/bin/env python
"""Bring-up probe for the channels-last slab kernel (conv_nhwc.cu).  Each case runs in its own subprocess (a trap
poisons the CUDA context) with barrier time-outs recorded instead of trapped.  Per case: error of FPROP / DGRAD against
torch fp64 with the descriptor base-offset field off (dbg 0, the product setting) and on (dbg 1), per-tap errors with one-hot filters (which
slab offsets work), and the time against the NCHW kernels.
  python tools/nhwc_probe.py                 -> all cases
  python tools/nhwc_probe.py case <i>        (internal)"""
import json
import os
import subprocess
import sys
import ctypes as C
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CASES = [  # B, Ci, H, W, Co, k, s, p
    (1, 32, 16, 8, 32, 3, 1, 1),          # one tile, one channel block
    (2, 64, 32, 40, 64, 3, 1, 1),
    (2, 32, 48, 72, 32, 7, 1, 3),
    (2, 196, 20, 44, 128, 3, 1, 1),
    (1, 512, 16, 24, 200, 3, 1, 1),
    (2, 64, 48, 64, 48, 3, 2, 1),
    (4, 64, 64, 208, 64, 3, 1, 1),        # real layers from here on (timed)
    (4, 128, 32, 104, 128, 3, 1, 1),
    (4, 256, 16, 52, 256, 3, 1, 1),
    (4, 32, 128, 416, 32, 7, 1, 3),
    (4, 196, 64, 208, 128, 3, 1, 1),
    (4, 128, 64, 208, 128, 3, 1, 1),
    (4, 128, 64, 208, 96, 3, 1, 1),
    (4, 96, 64, 208, 64, 3, 1, 1),
    (4, 64, 64, 208, 32, 3, 1, 1),
    (4, 129, 64, 208, 64, 3, 1, 1),
    (4, 65, 128, 416, 32, 3, 1, 1),
    (4, 16, 256, 832, 16, 3, 1, 1),       # 17: thin layers (dbg bit 1 routes them here instead of the direct kernel)
    (4, 17, 256, 832, 16, 3, 1, 1),
    (4, 16, 128, 416, 16, 3, 1, 1),
    (4, 32, 128, 416, 16, 3, 1, 1),
    (4, 32, 64, 208, 64, 3, 2, 1),        # 21: stride 2: dgrad classes
    (4, 16, 128, 416, 32, 3, 2, 1),
    (4, 512, 8, 26, 512, 3, 1, 1),        # 23: small maps (need CCB_NHWC_MINH=8 CCB_NHWC_WASTE10=30 to reach the kernel)
    (4, 1024, 8, 26, 512, 3, 1, 1),
    (4, 256, 8, 26, 256, 3, 1, 1),
    (4, 128, 8, 26, 128, 3, 1, 1),
]


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def run_case(i):
    from cc_b200 import nn as cnn, _lib
    import torch.nn.functional as F
    lib = _lib.lib()
    B, Ci, H, W, Co, k, s, p = CASES[i]
    dev = torch.device('cuda:0')
    g = torch.Generator().manual_seed(i)
    x = torch.randn(B, Ci, H, W, generator=g).to(dev).requires_grad_(True)
    w = (torch.randn(Co, Ci, k, k, generator=g) / (Ci * k * k) ** 0.5).to(dev).requires_grad_(True)
    b = torch.randn(Co, generator=g).to(dev)
    cnn.CONV_IMPL = _lib.IMPL_TC
    out = dict(case=list(CASES[i]))
    zd = F.conv2d(x.detach().double(), w.detach().double(), b.double(), s, p)
    wt = torch.randn(zd.shape, generator=g).to(dev)
    gxd = torch.autograd.grad((F.conv2d(x.detach().double().requires_grad_(True), w.detach().double(), b.double(), s, p) * wt.double()).sum(), [])\
        if False else None
    xd = x.detach().double().requires_grad_(True)
    wd = w.detach().double().requires_grad_(True)
    gxd, gwd = torch.autograd.grad((F.conv2d(xd, wd, b.double(), s, p) * wt.double()).sum(), [xd, wd])
    st = (C.c_uint * 4)()
    for dbg in (0, 2):
        lib.ccb_debug_nhwc(1, 1, dbg)
        y = cnn.conv2d(x, w, b, None, s, p, None, 0.0)
        gx, gw = torch.autograd.grad((y * wt).sum(), [x, w])
        torch.cuda.synchronize()
        lib.ccb_debug_nhwc_status(st)
        out['dbg%d' % (0 if dbg == 0 else 1)] = dict(fprop=rel(y.detach(), zd), dgrad=rel(gx, gxd), wgrad=rel(gw, gwd), status=list(st))
    # per-tap: one-hot filters (small cases only)
    if B * H * W <= 8192 and s == 1:
        taps = {}
        for dbg in (0,):
            lib.ccb_debug_nhwc(1, 1, dbg)
            errs = []
            for t in range(k * k):
                w1 = torch.zeros_like(w)
                w1.view(Co, Ci, k * k)[:, :, t] = w.detach().view(Co, Ci, k * k)[:, :, t]
                y = cnn.conv2d(x.detach(), w1, None, None, s, p, None, 0.0)
                errs.append(round(rel(y, F.conv2d(x.detach().double(), w1.double(), None, s, p)), 7))
            taps['dbg%d' % dbg] = errs
        out['tap_err'] = taps
    # timing: channels-last vs NCHW kernels (fprop and dgrad), 20 launches each
    def timed(fn):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 20 * 1e3
    gflop = 2.0 * B * zd.shape[2] * zd.shape[3] * Co * Ci * k * k / 1e9
    gy = wt
    for name, on, dbg in (('auto', 1, 0), ('nchw', 0, 0)):
        lib.ccb_debug_nhwc(on, 1, dbg)
        with torch.no_grad():
            tf = timed(lambda: cnn.conv2d(x, w, b, None, s, p, 'relu', 0.0))
        y = cnn.conv2d(x, w, b, None, s, p, None, 0.0)
        tb = timed(lambda: torch.autograd.grad((y,), [x], [gy], retain_graph=True))
        tw = timed(lambda: torch.autograd.grad((y,), [w], [gy], retain_graph=True))
        out['time_' + name] = dict(fprop_us=round(tf, 1), fprop_tflops=round(gflop / tf * 1e3, 1), dgrad_us=round(tb, 1),
                                   dgrad_tflops=round(gflop / tb * 1e3, 1), wgrad_us=round(tw, 1), wgrad_tflops=round(gflop / tw * 1e3, 1))
    lib.ccb_debug_nhwc(1, 1, 0)
    print(json.dumps(out))


if __name__ == '__main__':
    if len(sys.argv) >= 3 and sys.argv[1] == 'case':
        run_case(int(sys.argv[2]))
    else:
        sel = [int(a) for a in sys.argv[1:]] or range(len(CASES))
        for i in sel:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), 'case', str(i)], capture_output=True, text=True, timeout=300)
            line = [l for l in r.stdout.splitlines() if l.startswith('{')]
            print(line[-1] if line else json.dumps(dict(case=list(CASES[i]), rc=r.returncode, err=r.stderr[-400:])), flush=True)
