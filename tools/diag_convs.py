#!/usr/bin/env python
"""Per-call precision of the tensor-core conv kernels ON THE REAL DATA of a training step: the step runs on the exact
fp32 CUDA-core kernels (IMPL_FFMA); every conv call (fprop / dgrad / wgrad) is repeated on the tensor-core dispatch with
the channels-last kernel on and off, and the outputs are compared with the FFMA result (max-abs error relative to max-abs)."""
import ctypes as C
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.util import rel_err                      # noqa: E402
from cc_b200 import synth, nn as cnn, _lib          # noqa: E402
from cc_b200.train_step import Trainer              # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else 'cfg1'
B, H, W = 4, 256, 832
dev = torch.device('cuda:0')
tgt, refs = synth.frames(B, H, W, seed=50)
K, Kinv = synth.intrinsics(B, H, W)
lib = _lib.lib()
rows = []
orig_run = cnn._run
OUT_IDX = {0: 4, 1: 4, 2: 2}


def run_both(op, d, *args):
    orig_run(op, d, *args)                                          # IMPL_FFMA (cnn.CONV_IMPL at desc creation)
    ref = args[OUT_IDX[op]]
    ref_db = args[3] if op == 2 else None
    errs = {}
    for name, nhwc in (('nchw', 0), ('nhwc', 1)):
        lib.ccb_debug_nhwc(nhwc, 0, 0)
        d2 = _lib.ConvDesc()
        C.memmove(C.byref(d2), C.byref(d), C.sizeof(d))
        d2.impl = _lib.IMPL_TC
        d2.wcache = None
        a2 = list(args)
        a2[OUT_IDX[op]] = torch.empty_like(ref)
        if op == 2 and ref_db is not None:
            a2[3] = torch.empty_like(ref_db)
        orig_run(op, d2, *a2)
        errs[name] = rel_err(a2[OUT_IDX[op]], ref)
    lib.ccb_debug_nhwc(1, 0, 0)
    rows.append((('fprop', 'dgrad', 'wgrad')[op], (d.B, d.Ci, d.Hi, d.Wi, d.Co, d.kh, d.stride), errs['nchw'], errs['nhwc'],
                 float(ref.abs().max())))


cnn.CONV_IMPL = _lib.IMPL_FFMA
tr = Trainer(cfg, dev)
tr.wcache = None
cnn._run = run_both
tr.step(tgt.to(dev), [r.to(dev) for r in refs], K.to(dev), Kinv.to(dev))
cnn._run = orig_run
torch.cuda.synchronize()
rows.sort(key=lambda r: -max(r[2], r[3]))
print('%d conv calls; worst 40 (rel err of the tensor-core result against the fp32 CUDA-core result):' % len(rows))
for r in rows[:40]:
    print('%-6s B%d Ci%-4d %3dx%-3d Co%-4d k%d s%d   nchw %.2e  nhwc %.2e   |ref|max %.2e' % ((r[0],) + r[1] + (r[2], r[3], r[4])))
import statistics
for op in ('fprop', 'dgrad', 'wgrad'):
    e1 = [r[2] for r in rows if r[0] == op]
    e2 = [r[3] for r in rows if r[0] == op]
    print(op, len(e1), 'calls: nchw median %.1e max %.1e | nhwc median %.1e max %.1e' % (statistics.median(e1), max(e1), statistics.median(e2), max(e2)))
