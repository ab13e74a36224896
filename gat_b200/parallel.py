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


def share_seed():
    """make every rank use rank 0's placement seed (an unseeded run draws it from numpy's global generator,
    which differs between processes)"""
    import torch.distributed as dist
    from . import engine
    rank, world = rank_world()
    if world == 1:
        return engine.getSeed()
    box = [engine.getSeed() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    if rank != 0:
        calls = engine._rng_state["calls"]
        engine.seed(box[0])
        engine._rng_state["calls"] = calls
    return box[0]


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


def column_range(n_columns, rank, world):
    """[begin, end) of the annotation columns whose statistics `rank` computes (exchange_columns)"""
    return (rank * n_columns) // world, ((rank + 1) * n_columns) // world


def exchange_columns(local, num_samples):
    """The alternative to allgather_samples for very large S: instead of the whole S x A matrix
    on every rank, rank g ends up with ALL samples of its own columns column_range(A, g, G) -- 1/G of the
    all-gather's traffic and memory, and the column statistics (which are independent per column) run on A/G
    columns per GPU.  `local` = this rank's [shard_range(num_samples) rows][A] slab; returns [num_samples][A_g].
    One all_to_all_single with uneven splits (NCCL on GPUs, gloo in the CPU test).  gat_b200.run() takes this
    route (`exchange="columns"`, or automatically once the gathered matrices would exceed 1 GiB) instead of the
    all-gather north_star names: 1e6 samples x 1000 tracks are 4 GB per rank gathered, 0.5 GB column-sharded."""
    import torch.distributed as dist
    rank, world = rank_world()
    if world == 1:
        return local
    A = local.shape[1]
    rows = [shard_range(num_samples, r, world) for r in range(world)]
    cols = [column_range(A, r, world) for r in range(world)]
    n_mine = rows[rank][1] - rows[rank][0]
    a_mine = cols[rank][1] - cols[rank][0]
    assert local.shape[0] == n_mine, "slab does not match this rank's shard"
    # send buffer: for every destination rank its columns of my rows, row-major, back to back
    send = torch.cat([local[:, b:e].reshape(-1) for b, e in cols])
    in_splits = [n_mine * (e - b) for b, e in cols]
    out_splits = [(e - b) * a_mine for b, e in rows]
    recv = torch.empty(sum(out_splits), dtype=local.dtype, device=local.device)
    dist.all_to_all_single(recv, send, output_split_sizes=out_splits, input_split_sizes=in_splits)
    # pieces arrive in source-rank order = global sample order
    return recv.view(num_samples, a_mine) if a_mine else recv.view(num_samples, 0)


def allgather_columns(local, n_columns):
    """inverse of column_range() for small per-column results: `local` = [..., A_g] (this rank's columns, last
    dimension) -> [..., A] on every rank.  One all_gather of equal-sized (padded) pieces."""
    import torch.distributed as dist
    rank, world = rank_world()
    if world == 1:
        return local
    widths = [column_range(n_columns, r, world) for r in range(world)]
    widths = [e - b for b, e in widths]
    widest = max(widths)
    x = local
    if x.shape[-1] < widest:
        pad = torch.zeros(tuple(x.shape[:-1]) + (widest - x.shape[-1],), dtype=x.dtype, device=x.device)
        x = torch.cat([x, pad], dim=-1)
    pieces = [torch.empty_like(x) for _ in range(world)]
    dist.all_gather(pieces, x.contiguous())
    return torch.cat([pieces[r][..., :widths[r]] for r in range(world)], dim=-1)


class PeerUnavailable(RuntimeError):
    """the GPUs of this job cannot map each other's memory (no peer access / IPC): use the NCCL transport"""


class PeerMatrix(object):
    """One uint32 matrix [planes][rows][cols] per rank in memory that every other rank of the box can write
    (CUDA IPC over NVLink / NVSwitch): the destination of the counting kernel's output routes, i.e. of an exchange
    that needs no collective.  Collective constructor: every rank calls it with ITS shape; the IPC handles travel
    through torch.distributed's object collectives (control plane only).

        mine.tensor          this rank's matrix as a torch tensor (a view of the shared allocation)
        mine.pointer(r)      device pointer of rank r's matrix as mapped into this process
        mine.shape_of(r)     its (planes, rows, cols)
    """

    def __init__(self, ctx, planes, rows, cols):
        import torch.distributed as dist
        self.ctx = ctx
        self.rank, self.world = rank_world()
        self.shape = (int(planes), int(rows), int(cols))
        self.ptrs = None
        nbytes = max(4 * planes * rows * cols, 4)
        # every step is agreed on by all ranks, so that a GPU pair without peer access makes ALL of them raise
        # PeerUnavailable (and take the NCCL route) instead of leaving some waiting in a collective
        try:
            self.local_ptr, handle = ctx.peer_alloc(nbytes)
        except Exception as e:      # noqa: BLE001
            self.local_ptr, handle = None, None
            why = str(e)
        everyone = [None] * self.world
        if self.world > 1:
            dist.all_gather_object(everyone, (handle, self.shape))
        else:
            everyone[0] = (handle, self.shape)
        if any(h is None for h, _ in everyone):
            if self.local_ptr is not None:
                ctx.peer_free(self.local_ptr)
            raise PeerUnavailable("shareable device memory could not be allocated on every rank" +
                                  (": " + why if handle is None else ""))
        self.shapes = [s for _, s in everyone]
        ptrs, why = [], None
        for r, (h, _) in enumerate(everyone):
            try:
                ptrs.append(self.local_ptr if r == self.rank else ctx.peer_open(h))
            except Exception as e:  # noqa: BLE001
                why = str(e)
                break
        oks = [None] * self.world
        if self.world > 1:
            dist.all_gather_object(oks, why is None)
        else:
            oks[0] = why is None
        if not all(oks):
            for r, p in enumerate(ptrs):
                if r != self.rank:
                    ctx.peer_close(p)
            if self.world > 1:
                dist.all_gather_object(oks, True)      # (nobody frees memory a peer still has mapped)
            ctx.peer_free(self.local_ptr)
            raise PeerUnavailable("peer memory could not be mapped between all GPUs" + (": " + why if why else ""))
        self.ptrs = ptrs
        self.tensor = self._as_tensor()

    def _as_tensor(self):
        planes, rows, cols = self.shape

        class _Shared(object):      # the CUDA array interface lets torch view memory it did not allocate
            pass
        holder = _Shared()
        holder.__cuda_array_interface__ = {"shape": (planes, rows, cols), "typestr": "<i4", "data": (self.local_ptr, False),
                                           "version": 3, "strides": None}
        self._holder = holder
        return torch.as_tensor(holder, device=torch.device("cuda", self.ctx.device))

    def pointer(self, r):
        return self.ptrs[r]

    def shape_of(self, r):
        return self.shapes[r]

    def close(self):
        """collective in effect: call it on every rank once nobody writes any more (after a barrier)"""
        if self.ptrs is None:
            return
        self.tensor = None
        for r, p in enumerate(self.ptrs):
            if r != self.rank:
                self.ctx.peer_close(p)
        self.ctx.peer_free(self.local_ptr)
        self.ptrs = None


def barrier():
    """all ranks reach this point (and their GPUs are idle): after it, what the other ranks' kernels stored into this
    rank's PeerMatrix is complete"""
    import torch.distributed as dist
    rank, world = rank_world()
    on_gpu = torch.cuda.is_available()           # (the gloo-only CPU tests of the host logic have none)
    if on_gpu:
        torch.cuda.synchronize()
    if world > 1:
        box = [None] * world
        dist.all_gather_object(box, rank)        # (object collective: rides on the CPU backend when there is one)
    if on_gpu:
        torch.cuda.synchronize()


_warm = {"thread": None}


def warm_up_async():
    """Start NCCL's one-time set-up -- communicator creation and the lazily connected channels of the all-gather and
    of the all-to-all, several seconds on 8 GPUs -- on a background thread, so that it overlaps input parsing
    instead of sitting in front of the first exchange.  join_warm_up() waits for it (run() does, before its
    first collective)."""
    import threading
    import torch.distributed as dist
    rank, world = rank_world()
    if world == 1 or _warm["thread"] is not None or "nccl" not in str(dist.get_backend()):
        return

    def go():
        import os
        dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
        torch.cuda.set_device(dev)
        x = torch.zeros(world, dtype=torch.int32, device=dev)
        y = torch.empty(world * world, dtype=torch.int32, device=dev)
        dist.all_gather_into_tensor(y, x)
        dist.all_to_all_single(torch.empty_like(x), x)
        dist.all_gather([torch.empty_like(x) for _ in range(world)], x)
        torch.cuda.synchronize(dev)

    t = threading.Thread(target=go, name="gat_b200-nccl-warm-up", daemon=True)
    _warm["thread"] = t
    t.start()


def join_warm_up():
    t = _warm["thread"]
    if t is not None and t.is_alive():
        t.join()


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
        # NCCL for CUDA tensors, gloo for the small host-side object collectives (IPC handles, barriers, the seed):
        # a run that exchanges its counts through peer memory then never pays for NCCL's set-up
        backend = "cpu:gloo,cuda:nccl" if torch.cuda.is_available() else "gloo"
    if "nccl" in backend:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group(backend=backend)
    return True


def finalize():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        barrier()
        dist.destroy_process_group()
