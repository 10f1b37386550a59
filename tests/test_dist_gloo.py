"""The N>1 path of cc_b200.dist on CPU: world_size-2 gloo, flat-gradient all-reduce + averaged Adam step.
(The kernels run through the CPU simulator build; what is under test is the host-side exchange logic.)"""
import os
import sys
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests', 'sim'))
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR='127.0.0.1',
                      MASTER_PORT=str(port))
    import build_sim
    from cc_b200 import _lib, dist as cdist, nn as cnn
    from cc_b200.optim import FlatAdam
    _lib.use_library(build_sim.build())
    r, _, w = cdist.init_from_env(backend='gloo')
    assert (r, w) == (rank, world)
    torch.manual_seed(0)                                   # identical init on every rank
    conv = cnn.Conv2d(3, 4, 3, padding=1)
    opt = FlatAdam(conv.parameters(), lr=1e-2)
    cdist.broadcast_params(opt)
    x = torch.randn(2, 3, 6, 7, generator=torch.Generator().manual_seed(100 + rank))   # different shard per rank
    opt.zero_grad()
    (conv(x) ** 2).mean().backward()
    local_g = opt.flat_g.clone()
    cdist.allreduce_grads(opt)
    summed = opt.flat_g.clone()
    opt.step()
    ret[rank] = (local_g, summed, opt.flat_p.clone(), opt.grad_scale)
    cdist.barrier()
    torch.distributed.destroy_process_group()


def test_two_rank_allreduce_and_step():
    sys.path.insert(0, os.path.join(ROOT, 'tests', 'sim'))
    import build_sim
    build_sim.build()                       # build once in the parent so the workers only load it
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    g0, s0, p0, sc0 = ret[0]
    g1, s1, p1, sc1 = ret[1]
    assert sc0 == sc1 == 0.5
    assert torch.allclose(s0, g0 + g1, rtol=1e-6, atol=1e-8) and torch.equal(s0, s1)
    assert torch.equal(p0, p1)              # identical parameters after the averaged step
    # equals a single-process Adam step on the mean gradient
    sys.path.insert(0, ROOT)
    from oracle.step import Adam
    from cc_b200 import nn as cnn
    torch.manual_seed(0)
    conv = cnn.Conv2d(3, 4, 3, padding=1)
    flat = torch.cat([q.detach().reshape(-1) for q in conv.parameters()]).clone().requires_grad_(True)
    flat.grad = (g0 + g1) / 2
    o = Adam([flat], 1e-2)
    o.step()
    assert torch.allclose(flat.detach(), p0, rtol=1e-5, atol=1e-7)
