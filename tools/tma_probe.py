#!/usr/bin/env python
"""Bring-up probe for the TMA-fed tcgen05 conv kernel (conv_tma.cu).  Each case runs in its own subprocess (a trap
poisons the CUDA context); it compares TMA (debug flag 0) and register-gather (flag 8) tensor-core results with a
float64 torch reference and times both.
  python tools/tma_probe.py                 -> all cases
  python tools/tma_probe.py case <i>        (internal)"""
import json
import os
import subprocess
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CASES = [  # B, Ci, H, W, Co, k, s, p, op
    (1, 8, 4, 32, 16, 1, 1, 0, 'fprop'),         # direct TMA: one tile, one k-stage
    (1, 64, 8, 64, 128, 1, 1, 0, 'fprop'),       # direct TMA: 2 k-stages, N = 128
    (1, 8, 4, 32, 16, 3, 1, 1, 'fprop'),         # slab: 9 taps, one slab, padding through OOB fill
    (2, 32, 16, 32, 64, 3, 1, 1, 'fprop'),       # slab: one block, 9 k-stages
    (2, 20, 13, 36, 40, 3, 1, 1, 'fprop'),       # ragged: 20 channels (K = 180 -> 6 stages), row / column tails
    (2, 12, 10, 44, 200, 5, 1, 2, 'fprop'),      # N > 128 (two n-tiles), 5x5
    (2, 3, 20, 64, 32, 7, 2, 3, 'fprop'),        # stride 2, 3 channels, 7x7 (disp conv1)
    (2, 32, 16, 64, 64, 3, 2, 1, 'fprop'),       # stride 2
    (2, 160, 8, 32, 48, 3, 1, 1, 'fprop'),       # several channel blocks (double-buffered slabs)
    (2, 32, 16, 64, 256, 3, 2, 1, 'dgrad'),      # 1-tap parity class with many short channel blocks
    (2, 64, 32, 64, 128, 3, 2, 1, 'dgrad'),
    (2, 32, 16, 64, 64, 1, 2, 0, 'fprop'),       # 1x1 stride 2
    (4, 128, 32, 104, 128, 3, 1, 1, 'fprop'),    # real layers from here on
    (4, 32, 128, 416, 32, 7, 1, 3, 'fprop'),
    (4, 16, 256, 832, 16, 3, 1, 1, 'fprop'),
    (4, 17, 256, 832, 16, 3, 1, 1, 'fprop'),
    (4, 32, 128, 416, 64, 3, 2, 1, 'fprop'),
    (4, 15, 256, 832, 16, 7, 2, 3, 'fprop'),
    (2, 32, 16, 32, 64, 3, 1, 1, 'dgrad'),
    (2, 24, 32, 64, 48, 3, 2, 1, 'dgrad'),       # stride-2 parity classes
    (2, 24, 8, 32, 12, 4, 2, 1, 'convT'),        # ConvTranspose2d forward (Back2Future upsampling)
    (4, 128, 32, 104, 128, 3, 1, 1, 'dgrad'),
    (4, 32, 128, 416, 32, 7, 1, 3, 'dgrad'),
    (4, 32, 256, 832, 64, 3, 2, 1, 'dgrad'),     # stride-2 layer at full size
    (1, 8, 4, 32, 16, 1, 1, 0, 'wgrad'),         # one M tile, one tap
    (2, 16, 8, 64, 32, 3, 1, 1, 'wgrad'),        # two tap groups (5 + 4 taps), padding
    (2, 40, 9, 44, 136, 3, 1, 1, 'wgrad'),       # ragged: 40 channels (3 taps per tile), N > 128, row tail
    (2, 24, 16, 64, 16, 3, 2, 1, 'wgrad'),       # stride 2
    (2, 160, 8, 32, 48, 3, 1, 1, 'wgrad'),       # channel blocks
    (4, 32, 128, 416, 32, 7, 1, 3, 'wgrad'),
    (4, 16, 256, 832, 16, 3, 1, 1, 'wgrad'),
    (4, 16, 256, 832, 1, 3, 1, 1, 'wgrad'),
    (4, 128, 32, 104, 128, 3, 1, 1, 'wgrad'),
    (4, 32, 128, 416, 64, 3, 2, 1, 'wgrad'),
    (4, 16, 250, 832, 16, 3, 1, 1, 'fprop'),     # stacked tiles with a ragged bottom
    (4, 16, 250, 832, 16, 3, 1, 1, 'dgrad'),
    (4, 16, 256, 832, 1, 3, 1, 1, 'fprop'),      # 36: thin layers for the direct kernel
    (4, 3, 256, 832, 32, 7, 2, 3, 'fprop'),
    (4, 17, 256, 832, 16, 1, 1, 0, 'fprop'),
    (4, 32, 128, 416, 16, 3, 2, 1, 'convT'),
    (4, 15, 256, 832, 16, 7, 2, 3, 'dgrad'),
    (4, 65, 128, 416, 32, 3, 1, 1, 'dgrad'),
    (4, 65, 128, 416, 32, 3, 1, 1, 'fprop'),
    (4, 32, 128, 416, 32, 3, 1, 1, 'fprop'),
    (4, 16, 128, 416, 32, 5, 2, 2, 'fprop'),
    (2, 13, 70, 100, 20, 3, 1, 1, 'fprop'),      # 45: ragged everything (direct kernel needs out_px >= 148K: not taken)
    (6, 13, 100, 260, 20, 3, 1, 1, 'fprop'),     # 46: ragged, large enough for the direct kernel
    (6, 13, 100, 260, 20, 3, 2, 1, 'dgrad'),
    (4, 512, 8, 26, 512, 3, 1, 1, 'fprop'),      # 48: narrow maps (W = 26): padded rows
    (4, 512, 8, 26, 512, 3, 1, 1, 'dgrad'),
    (4, 512, 8, 26, 512, 3, 1, 1, 'wgrad'),
    (4, 256, 16, 52, 512, 3, 2, 1, 'dgrad'),
    (4, 256, 16, 52, 512, 3, 2, 1, 'wgrad'),
    (2, 24, 9, 22, 40, 3, 1, 1, 'fprop'),        # 53: small ragged narrow map
    (2, 24, 9, 22, 40, 3, 1, 1, 'dgrad'),
    (2, 24, 9, 22, 40, 3, 1, 1, 'wgrad'),
]


if os.environ.get('CCB_PROBE_ALT'):
    CASES[:] = [(1, 8, 6, 32, 16, 1, 1, 0, 'fprop'), (1, 8, 4, 32, 16, 3, 1, 1, 'fprop')]


def run_case(i):
    from cc_b200 import nn as cnn, _lib
    import torch.nn.functional as F
    B, Ci, H, W, Co, k, s, p, op = CASES[i]
    dev = torch.device('cuda:0')
    g = torch.Generator().manual_seed(i)
    out = {}
    x0 = torch.randn(B, Ci, H, W, generator=g)
    wshape = (Ci, Co, k, k) if op == 'convT' else (Co, Ci, k, k)
    w0 = torch.randn(wshape, generator=g) * 0.1
    b0 = torch.randn(Co, generator=g)
    for flag, impl, name in ((8, _lib.IMPL_TC, 'gather3'), (16, _lib.IMPL_TC_TF32, 'tma1'), (16, _lib.IMPL_TC, 'tma3'),
                             (8, _lib.IMPL_TC_TF32, 'gather1')):
        _lib.lib().ccb_debug_tc_swap_strides(flag)
        cnn.CONV_IMPL = impl
        x = x0.to(dev).requires_grad_(True)
        w = w0.to(dev).requires_grad_(True)
        b = b0.to(dev)
        if op == 'convT':
            fn = lambda: cnn.conv_transpose2d(x, w, b, s, p, 0, 'relu')
            res = fn()
            ref = F.relu(F.conv_transpose2d(x.double(), w.double(), b.double(), s, p, 0))
        elif op == 'fprop':
            fn = lambda: cnn.conv2d(x, w, b, None, s, p, 'leaky', 0.2)
            res = fn()
            ref = F.leaky_relu(F.conv2d(x.double(), w.double(), b.double(), s, p), 0.2)
        else:
            y = cnn.conv2d(x, w, b, None, s, p, None, 0.0)
            g.manual_seed(100 + i)
            wt = torch.randn(y.shape, generator=g).to(dev)
            wrt = x if op == 'dgrad' else w
            fn = lambda: torch.autograd.grad((y * wt).sum(), [wrt], retain_graph=True)[0]
            res = fn()
            xd = x.detach().double().requires_grad_(True)
            wd = w.detach().double().requires_grad_(True)
            yd2 = F.conv2d(xd, wd, b.double(), s, p)
            ref = torch.autograd.grad((yd2 * wt.double()).sum(), [xd if op == 'dgrad' else wd])[0]
        torch.cuda.synchronize()
        out[name] = ((res.double() - ref).abs().max() / ref.abs().max()).item()
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        out[name + '_ms'] = round(e0.elapsed_time(e1) / 10, 4)
        if name.startswith('tma'):
            import ctypes
            st = (ctypes.c_uint * 4)()
            _lib.lib().ccb_debug_tma_status(st)
            if st[0]:
                out[name + '_timeout'] = list(st)
    print(json.dumps(dict(case=CASES[i], **out)))


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'case':
        run_case(int(sys.argv[2]))
    else:
        sel = [int(v) for v in os.environ.get('CCB_PROBE_CASES', '').split(',') if v] or range(len(CASES))
        for i in sel:
            try:
                r = subprocess.run([sys.executable, __file__, 'case', str(i)], capture_output=True, text=True, timeout=120)
                line = [l for l in r.stdout.splitlines() if l.startswith('{')]
                print(line[-1] if line else json.dumps(dict(case=CASES[i], rc=r.returncode,
                                                            err=(r.stderr.strip().splitlines() or ['?'])[-3:])))
            except subprocess.TimeoutExpired:
                print(json.dumps(dict(case=CASES[i], timeout=True)))
            sys.stdout.flush()
