#!/usr/bin/env python
"""Bring-up probe for the tcgen05 conv kernel: each case runs in its own subprocess (a trap poisons
the CUDA context) and compares the tensor-core result with the FFMA kernel's.
  python tools/tc_probe.py            -> runs all cases
  python tools/tc_probe.py case <i> <swap>   (internal)"""
import json
import os
import subprocess
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CASES = [  # B, Ci, H, W, Co, k, s, p, op
    (1, 8, 8, 16, 16, 1, 1, 0, 'fprop'),        # M=128, K=8 -> one k-tile, N=16: the minimal MMA
    (1, 32, 8, 16, 16, 1, 1, 0, 'fprop'),       # one full k-tile
    (1, 64, 8, 16, 128, 1, 1, 0, 'fprop'),      # 2 k-tiles, N=128
    (2, 32, 16, 24, 64, 3, 1, 1, 'fprop'),      # taps + padding, M tail
    (2, 17, 13, 19, 40, 3, 2, 1, 'fprop'),      # odd everything, stride 2
    (4, 128, 32, 104, 128, 3, 1, 1, 'fprop'),   # real layer
    (4, 32, 128, 416, 32, 7, 1, 3, 'fprop'),    # heaviest disp layer
    (2, 32, 16, 24, 64, 3, 1, 1, 'dgrad'),
    (2, 48, 9, 14, 24, 3, 2, 1, 'dgrad'),
    (2, 24, 4, 6, 12, 4, 2, 1, 'convT'),
]


def run_case(i, swap):
    from cc_b200 import nn as cnn, _lib
    import torch.nn.functional as F
    B, Ci, H, W, Co, k, s, p, op = CASES[i]
    dev = torch.device('cuda:0')
    g = torch.Generator().manual_seed(i)
    _lib.lib().ccb_debug_tc_swap_strides(swap)
    out = {}
    for impl, name in ((_lib.IMPL_FFMA, 'ffma'), (_lib.IMPL_TC, 'tc3'), (_lib.IMPL_TC_TF32, 'tc1')):
        cnn.CONV_IMPL = impl
        g.manual_seed(i)
        x = torch.randn(B, Ci, H, W, generator=g).to(dev).requires_grad_(True)
        if op == 'convT':
            w = (torch.randn(Ci, Co, k, k, generator=g) * 0.1).to(dev).requires_grad_(True)
            b = torch.randn(Co, generator=g).to(dev)
            y = cnn.conv_transpose2d(x, w, b, s, p, 0, 'relu')
            ref = F.relu(F.conv_transpose2d(x.double(), w.double(), b.double(), s, p, 0))
            res = y
        else:
            w = (torch.randn(Co, Ci, k, k, generator=g) * 0.1).to(dev).requires_grad_(True)
            b = torch.randn(Co, generator=g).to(dev)
            y = cnn.conv2d(x, w, b, None, s, p, 'leaky', 0.2)
            yd = F.leaky_relu(F.conv2d(x.double(), w.double(), b.double(), s, p), 0.2)
            if op == 'fprop':
                res, ref = y, yd
            else:
                wt = torch.randn(y.shape, generator=g).to(dev)
                res = torch.autograd.grad((y * wt).sum(), [x])[0]
                xd = x.detach().double().requires_grad_(True)
                yd2 = F.leaky_relu(F.conv2d(xd, w.double(), b.double(), s, p), 0.2)
                ref = torch.autograd.grad((yd2 * wt.double()).sum(), [xd])[0]
        torch.cuda.synchronize()
        err = ((res.double() - ref).abs().max() / ref.abs().max()).item()
        out[name] = err
        # timing
        if op == 'fprop':
            fn = (lambda: cnn.conv2d(x, w, b, None, s, p, 'leaky', 0.2))
            for _ in range(3):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                fn()
            e1.record()
            torch.cuda.synchronize()
            out[name + '_ms'] = e0.elapsed_time(e1) / 10
    flops = 2.0 * B * Co * Ci * k * k * ((H + 2 * p - k) // s + 1) * ((W + 2 * p - k) // s + 1)
    out['gflop'] = flops / 1e9
    print(json.dumps(dict(case=CASES[i], swap=swap, **out)))


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'case':
        run_case(int(sys.argv[2]), int(sys.argv[3]))
    else:
        for swap in (0, 4):
            for i in range(len(CASES)):
                if swap == 4 and i not in (2, 5, 6):
                    continue
                try:
                    r = subprocess.run([sys.executable, __file__, 'case', str(i), str(swap)], capture_output=True,
                                       text=True, timeout=120)
                    line = [l for l in r.stdout.splitlines() if l.startswith('{')]
                    print(line[-1] if line else json.dumps(dict(case=CASES[i], swap=swap, rc=r.returncode,
                                                                err=r.stderr.strip().splitlines()[-3:])))
                except subprocess.TimeoutExpired:
                    print(json.dumps(dict(case=CASES[i], swap=swap, timeout=True)))
                sys.stdout.flush()
