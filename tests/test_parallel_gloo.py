"""world_size-2 gloo test of the sample sharding + single all-gather (gat_b200/parallel.py)."""
import os
import socket
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, %r)
    import torch, torch.distributed as dist
    from gat_b200 import parallel
    assert parallel.init_from_env(backend="gloo")
    rank, world = parallel.rank_world()
    assert world == 2
    for S in (10, 11, 1, 7):
        b, e = parallel.shard_range(S, rank, world)
        # every rank fills its shard of a [counters][samples][annotations] slab with the global sample id
        full = torch.arange(S, dtype=torch.int32).view(1, S, 1).expand(2, S, 3).contiguous()
        local = full[:, b:e, :].contiguous()
        got = parallel.allgather_samples(local, S, dim=1)
        assert got.shape == full.shape and torch.equal(got, full), (S, rank)
        got0 = parallel.allgather_samples(full[0, b:e, :].contiguous().double(), S, dim=0)
        assert torch.equal(got0, full[0].double())
    # all-to-all by annotation column: every rank ends up with all samples of its own columns
    for S, A in ((10, 6), (11, 5), (3, 1), (7, 9)):
        b, e = parallel.shard_range(S, rank, world)
        full = (torch.arange(S, dtype=torch.int32).view(S, 1) * 1000 + torch.arange(A, dtype=torch.int32).view(1, A))
        mine = parallel.exchange_columns(full[b:e].contiguous(), S)
        cb, ce = parallel.column_range(A, rank, world)
        assert mine.shape == (S, ce - cb) and torch.equal(mine, full[:, cb:ce]), (S, A, rank)
        # per-column results (statistics) of every rank's columns, gathered back to all columns
        stats = torch.arange(6 * A, dtype=torch.float64).view(6, A)
        back = parallel.allgather_columns(stats[:, cb:ce].contiguous(), A)
        assert torch.equal(back, stats), (A, rank)
    # an unseeded run: every rank ends up with rank 0's seed
    from gat_b200 import engine
    import numpy as np
    np.random.seed(1000 + rank)
    seed = parallel.share_seed()
    box = [seed]
    dist.broadcast_object_list(box, src=0)
    assert box[0] == seed == engine.getSeed(), (rank, seed, box)
    parallel.finalize()
    print("rank", rank, "ok")
""")


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_ranges_cover_everything():
    from gat_b200 import parallel
    for S in (0, 1, 7, 1000003):
        for world in (1, 2, 3, 8):
            r = [parallel.shard_range(S, g, world) for g in range(world)]
            assert r[0][0] == 0 and r[-1][1] == S
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            assert max(e - b for b, e in r) - min(e - b for b, e in r) <= 1


def test_allgather_two_ranks_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    port = free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=240)
        assert p.returncode == 0, out
        assert "ok" in out
