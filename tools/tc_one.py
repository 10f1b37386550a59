#!/usr/bin/env python
"""Run one convolution layer (fwd or fwd+bwd) a few times - target for `ncu --set full -k regex:conv_tc`."""
import argparse
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cc_b200 import nn as cnn, _lib   # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--shape', default='4,128,32,104,128,3,1,1')     # B,Ci,H,W,Co,k,s,p
ap.add_argument('--impl', type=int, default=_lib.IMPL_TC)
ap.add_argument('--bwd', action='store_true')
ap.add_argument('--iters', type=int, default=3)
a = ap.parse_args()
B, Ci, H, W, Co, k, s, p = [int(v) for v in a.shape.split(',')]
cnn.CONV_IMPL = a.impl
dev = torch.device('cuda:0')
x = torch.randn(B, Ci, H, W, device=dev, requires_grad=True)
w = (torch.randn(Co, Ci, k, k, device=dev) * 0.05).requires_grad_(True)
b = torch.zeros(Co, device=dev, requires_grad=True)
for _ in range(a.iters):
    y = cnn.conv2d(x, w, b, None, s, p, 'relu')
    if a.bwd:
        y.backward(torch.ones_like(y))
        x.grad = w.grad = b.grad = None
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    y = cnn.conv2d(x, w, b, None, s, p, 'relu')
    if a.bwd:
        y.backward(torch.ones_like(y))
        x.grad = w.grad = b.grad = None
e1.record()
torch.cuda.synchronize()
print('ms per iter', e0.elapsed_time(e1) / 10)
