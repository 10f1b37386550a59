"""Data-parallel plumbing: one process per GPU, one NCCL all-reduce of the flat gradient buffer per step.

Replaces the reference's four nn.DataParallel wrappers (train.py:300-303), which scatter the batch,
replicate the modules and gather outputs to GPU0 every step and run every loss on GPU0.  Here each
rank runs the whole step (nets + losses) on its own b/N shard - samples are independent (SURVEY.md 8e;
the batch-global oob normalisation and BatchNorm statistics are per-replica in the reference too) -
and the only exchange is ncclAllReduce(sum) over the flat fp32 gradient buffer, averaged by the Adam
kernel's grad_scale = 1/world_size."""
import os
import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """torchrun-style env (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*). Returns (rank, local_rank, world)."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29500')
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device('cuda', local))
        else:
            dist.init_process_group(backend)
    return rank, local, world


def bind_to_gpu_numa(local_rank):
    """Pin this process to the CPUs next to its GPU (NVML's ideal affinity) so that pinned host batches are allocated on
    the GPU's NUMA node: a cross-socket H2D copy runs at a fraction of the PCIe rate (seen on the 2-GPU box: GPU1 sits
    on NUMA node 1).  Best effort: any failure leaves the affinity unchanged."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get('CUDA_VISIBLE_DEVICES')
        idx = int(vis.split(',')[local_rank]) if vis and vis.split(',')[local_rank].isdigit() else local_rank
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(idx))
        return True
    except Exception:
        return False


def world_size():
    return dist.get_world_size() if dist.is_initialized() else 1


def allreduce_grads(opt):
    """Sum the flat gradient buffer across ranks (one collective) and let Adam average it."""
    w = world_size()
    if w > 1:
        dist.all_reduce(opt.flat_g, op=dist.ReduceOp.SUM)
    opt.grad_scale = 1.0 / w


def broadcast_params(opt, src=0):
    """Once at start-up (identical seeds make this a no-op in practice); never per step."""
    if world_size() > 1:
        dist.broadcast(opt.flat_p, src=src)


def max_over_ranks(value, device):
    t = torch.tensor([float(value)], device=device, dtype=torch.float64)
    if world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def barrier():
    if world_size() > 1:
        dist.barrier()
