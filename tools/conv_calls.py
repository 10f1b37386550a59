#!/usr/bin/env python
"""Every convolution call of one eager training step with the kernel it dispatched to and its CUDA-event time (helpers of
the call included): where the conv time of the step goes, per layer.  python tools/conv_calls.py [cfg] [top]"""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cc_b200 import synth, _lib, pyramid                 # noqa: E402
from cc_b200.train_step import Trainer                   # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else 'cfg3'
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
B, H, W = 4, 256, 832
dev = torch.device('cuda:0')
tgt, refs = synth.frames(B, H, W, seed=1)
K, Kinv = synth.intrinsics(B, H, W)
tr = Trainer(cfg, dev)
args = (tgt.to(dev), [r.to(dev) for r in refs], K.to(dev), Kinv.to(dev))
real = _lib.lib()
records = []


class Proxy:
    def __getattr__(self, name):
        fn = getattr(real, name)
        if not name.startswith('ccb_conv2d_'):
            return fn

        def wrapped(*a):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            rc = fn(*a)
            e.record()
            d = a[0]._obj
            records.append((name[len('ccb_conv2d_'):], (real.ccb_debug_last_conv_kernel() or b'').decode(),
                            (d.B, d.Ci, d.Hi, d.Wi, d.Co, d.kh, d.stride), s, e))
            return rc
        return wrapped


for it in range(3):
    records.clear()
    _lib._lib = Proxy() if it == 2 else real
    pyramid.clear()
    tr.step(*args)
    _lib._lib = real
    torch.cuda.synchronize()
rows = {}
for op, kern, shp, s, e in records:
    r = rows.setdefault((op, kern, shp), [0.0, 0])
    r[0] += s.elapsed_time(e)
    r[1] += 1
tot = sum(r[0] for r in rows.values())
print('%d conv calls, %.2f ms (eager, events per call)' % (len(records), tot))
by_k = {}
for (op, kern, shp), (ms, n) in rows.items():
    k = by_k.setdefault(kern + ':' + op, [0.0, 0, 0.0])
    k[0] += ms; k[1] += n
    k[2] += n * 2.0 * shp[0] * shp[4] * shp[1] * shp[5] ** 2 * (shp[2] // shp[6]) * (shp[3] // shp[6])
for k, (ms, n, fl) in sorted(by_k.items(), key=lambda kv: -kv[1][0]):
    print('  %-24s %6.2f ms %5.1f%%  %3d calls  %6.1f TFLOP/s' % (k, ms, 100 * ms / tot, n, fl / ms / 1e9))
print('top %d (op kernel B Ci HxW Co k s | calls ms us/call TFLOP/s):' % top)
for (op, kern, shp), (ms, n) in sorted(rows.items(), key=lambda kv: -kv[1][0])[:top]:
    fl = 2.0 * shp[0] * shp[4] * shp[1] * shp[5] ** 2 * (shp[2] // shp[6]) * (shp[3] // shp[6])
    print('  %-5s %-16s B%d Ci%-4d %3dx%-3d Co%-4d k%d s%d | %2d %6.2f %7.1f %6.1f' % ((op, kern) + shp + (n, ms, 1e3 * ms / n, n * fl / ms / 1e9)))
