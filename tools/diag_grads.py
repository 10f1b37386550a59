#!/usr/bin/env python
"""Which conv path is responsible for a gradient deviation?  cfg1 step at full size against the oracle (on the GPU: ATen
fp32, TF32 off; and on the CPU), with the product on: FFMA (exact fp32 CUDA cores), tensor cores with the channels-last
kernel off, tensor cores without the weight cache, the production dispatch."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.util import rel_err                      # noqa: E402
from tests import fullsize_cases as FC              # noqa: E402
from cc_b200 import synth, nn as cnn, _lib          # noqa: E402
from cc_b200.train_step import Trainer              # noqa: E402
from oracle import step as OS                       # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
cfg = sys.argv[1] if len(sys.argv) > 1 else 'cfg1'
B, H, W = 4, 256, 832
dev = torch.device('cuda:0')
tgt, refs = synth.frames(B, H, W, seed=50)
K, Kinv = synth.intrinsics(B, H, W)
P = OS.make_params(cfg)
ovals, ograds = FC._oracle(cfg, P, tgt, refs, K, Kinv)
sd = {n: {k: v.detach().clone() for k, v in d.items()} for n, d in P.items()}
dt, dr, dK, dKi = tgt.to(dev), [r.to(dev) for r in refs], K.to(dev), Kinv.to(dev)
fvals, fgrads = FC._oracle(cfg, FC._params_to(P, dev), dt, dr, dK, dKi)
lib = _lib.lib()
res = {}
for name, impl, nhwc, wcache in (('ffma', _lib.IMPL_FFMA, 1, True), ('tc_nchw', _lib.IMPL_AUTO, 0, True),
                                 ('auto_nocache', _lib.IMPL_AUTO, 1, False), ('auto', _lib.IMPL_AUTO, 1, True)):
    cnn.CONV_IMPL = impl
    lib.ccb_debug_nhwc(nhwc, 0, 0)
    tr = Trainer(cfg, dev, state_dicts=sd)
    if not wcache:
        tr.wcache = None
    tr.step(dt, dr, dK, dKi)
    g = {}
    for n in tr.nets:
        for k, p in tr.nets[n].named_parameters():
            if getattr(p, '_ccb_grad', None) is not None:
                g['%s.%s' % (n, k)] = p._ccb_grad.clone()
    res[name] = g
    del tr
lib.ccb_debug_nhwc(1, 0, 0)
names = sorted(ograds, key=lambda k: -rel_err(res['auto'][k], ograds[k]))[:25]
print('%-44s %9s | %9s %9s %9s %9s | vs GPU oracle: %9s %9s' % ('tensor', 'floor', 'ffma', 'tc_nchw', 'nocache', 'auto', 'ffma', 'auto'))
for k in names:
    print('%-44s %9.2e | %9.2e %9.2e %9.2e %9.2e | %9.2e %9.2e' % (
        k, rel_err(fgrads[k], ograds[k]), rel_err(res['ffma'][k], ograds[k]), rel_err(res['tc_nchw'][k], ograds[k]),
        rel_err(res['auto_nocache'][k], ograds[k]), rel_err(res['auto'][k], ograds[k]),
        rel_err(res['ffma'][k], fgrads[k]), rel_err(res['auto'][k], fgrads[k])))
def l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))
print('relative L2 error of the same tensors (floor | ffma | auto):')
for k in names[:12]:
    print('%-44s %9.2e | %9.2e %9.2e' % (k, l2(fgrads[k], ograds[k]), l2(res['ffma'][k], ograds[k]), l2(res['auto'][k], ograds[k])))
import statistics
for v in ('ffma', 'tc_nchw', 'auto_nocache', 'auto'):
    r = [rel_err(res[v][k], ograds[k]) / max(rel_err(fgrads[k], ograds[k]), 1e-7) for k in ograds]
    print(v, 'err/floor: median %.2f  p90 %.2f  max %.1f' % (statistics.median(r), sorted(r)[int(0.9 * len(r))], max(r)))
