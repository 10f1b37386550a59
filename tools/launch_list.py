#!/usr/bin/env python
"""One eager training step of a configuration inside a cudaProfilerStart/Stop range, for
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file X \
      python tools/launch_list.py --cfg cfg3
(the per-launch durations are cold-cache and serialised: compare SHARES).  `--summarise X` turns the csv
into the per-kernel table committed under profiles/."""
import argparse
import collections
import csv
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def summarise(path, out=sys.stdout):
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ik, iv, iu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    for r in rd:
        v = float(r[iv].replace(',', ''))
        v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'nsecond': 1e-3, 'usecond': 1.0, 'msecond': 1e3}.get(r[iu], 1e-3)
        name = re.sub(r'^void ', '', r[ik])
        name = re.sub(r'\(.*$', '', name)
        rows.append((name, v))
    tot = sum(v for _, v in rows)
    agg = collections.OrderedDict()
    for n, v in rows:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += v
    out.write('# %d launches, %.2f ms of kernel time (serialised, cold cache: compare SHARES)\n' % (len(rows), tot / 1e3))
    ours = sum(v for n, v in rows if not n.startswith('at::') and 'nccl' not in n.lower())
    out.write('# repo kernels: %.1f%% of kernel time; ATen glue: %.1f%%\n' % (100 * ours / tot, 100 * (tot - ours) / tot))
    for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.write('%-70s n=%4d sum=%9.1fus %5.1f%%\n' % (n[:70], c, v, 100 * v / tot))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--cfg', default='cfg3')
    ap.add_argument('--batch', type=int, default=4)
    ap.add_argument('--summarise')
    a = ap.parse_args()
    if a.summarise:
        return summarise(a.summarise)
    import torch
    from cc_b200 import pyramid, synth
    from cc_b200.train_step import Trainer
    dev = torch.device('cuda:0')
    tr = Trainer(a.cfg, dev, seed=0)
    tgt, refs = synth.frames(a.batch, 256, 832, seed=1)
    K, Kinv = synth.intrinsics(a.batch, 256, 832)
    tgt, refs, K, Kinv = tgt.to(dev), [r.to(dev) for r in refs], K.to(dev), Kinv.to(dev)
    for _ in range(2):
        pyramid.clear()
        tr.step(tgt, refs, K, Kinv)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    pyramid.clear()
    loss, _ = tr.step(tgt, refs, K, Kinv)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print('loss', float(loss))


if __name__ == '__main__':
    main()
