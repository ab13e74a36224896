"""Diagnostic: cost of creating / destroying an annotation set (stream-ordered allocations, 1.3 GB) before and after this
process has mapped a peer GPU's memory through CUDA IPC.  torchrun, 2 ranks."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group(backend="cpu:gloo,cuda:nccl")
    rank = dist.get_rank()
    from gat_b200 import device, parallel
    ctx = device.Context(local)
    rng = np.random.default_rng(0)
    A, K, n = 200, 24, 20000
    lists = []
    for a in range(A):
        row = []
        for k in range(K):
            s = np.sort(rng.choice(10 ** 7, size=n // K, replace=False)).astype(np.uint32) * 20
            row.append(np.stack([s, s + 10], axis=1).astype(np.uint32))
        lists.append(row)
    offs, start, end = device.to_csr([lists[a][k] for a in range(A) for k in range(K)])
    pin = [torch.from_numpy(x.view(np.int64 if x.dtype == np.uint64 else np.int32)).pin_memory() for x in (offs, start, end)]
    csr = (A, K, pin[0].numpy().view(np.uint64), pin[1].numpy().view(np.uint32), pin[2].numpy().view(np.uint32))

    def cycle(tag, reps=6):
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            an = device.Annotations(ctx, None, csr=csr, lazy=True)
            t1 = time.perf_counter()
            an.wait()
            t2 = time.perf_counter()
            an.close()
            torch.cuda.synchronize()
            t3 = time.perf_counter()
            ts.append((t1 - t0, t2 - t1, t3 - t2))
        if rank == 0:
            print(tag, " ".join("create %.2f wait %.2f close %.2f ms |" % tuple(1e3 * x for x in t) for t in ts))

    cycle("before peer mapping:")
    pm = parallel.PeerMatrix(ctx, 1, 1000, 1000)
    cycle("after peer mapping: ")
    parallel.barrier()
    pm.close()
    cycle("after closing it:   ")
    parallel.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
