"""Multi-GPU: samples are independent (gat/__init__.py:738-747), so rank g of G computes the global sample
indices [g*S/G, (g+1)*S/G) with every input replicated, and ONE collective (all-gather of the
S/G x A count slabs) rebuilds the S x A matrix on every rank.  The placement stream is keyed
by the GLOBAL sample index, so the matrix is bit-identical for any number of GPUs.

One process per GPU, launched with torch.distributed.run; NCCL over NVLink on GPUs, gloo in CPU tests.
"""
import torch


def rank_world():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(num_samples, rank, world):
    """[begin, end) of the global sample indices owned by `rank`"""
    return (rank * num_samples) // world, ((rank + 1) * num_samples) // world


def allgather_samples(local, num_samples, dim=0):
    """all-gather a tensor sharded along `dim` by shard_range(); returns the full tensor on every rank"""
    import torch.distributed as dist
    rank, world = rank_world()
    if world == 1:
        return local
    sizes = [shard_range(num_samples, r, world) for r in range(world)]
    sizes = [e - b for b, e in sizes]
    longest = max(sizes)
    x = local.movedim(dim, 0)
    x = x[:sizes[rank]].contiguous()
    if x.shape[0] < longest:                      # equal-sized slabs for one all_gather_into_tensor
        pad = torch.zeros((longest - x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        x = torch.cat([x, pad], dim=0)
    out = torch.empty((world * longest,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, x)
    if any(s != longest for s in sizes):
        out = torch.cat([out[r * longest:r * longest + sizes[r]] for r in range(world)], dim=0)
    return out.movedim(0, dim).contiguous()


def init_from_env(backend=None):
    """join the process group described by RANK/WORLD_SIZE/MASTER_* (set by torch.distributed.run);
    a plain single-process start does nothing."""
    import os
    import torch.distributed as dist
    if "RANK" not in os.environ or int(os.environ.get("WORLD_SIZE", "1")) <= 1:
        return False
    if dist.is_initialized():
        return True
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group(backend=backend)
    return True


def finalize():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()
