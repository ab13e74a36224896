"""Randomised counting parity beyond the test suite's budget: random annotation sets and segment lists, all
counters, compared bit for bit with the oracle -- verification aid

    python tools/stress_count.py [n_problems] [seed]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import helpers  # noqa: E402
from gat_b200 import device  # noqa: E402
from oracle import oracle  # noqa: E402


def make_context():
    """cuda:0, or with --emu the SIMT-emulated build of the kernels (tests/emu: no GPU needed, much slower)"""
    if EMU:
        sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
        import emu_context
        return emu_context.context()
    return device.Context(0)


EMU = "--emu" in sys.argv
if EMU:
    sys.argv.remove("--emu")

COUNTERS = ["nucleotide-overlap", "nucleotide-density", "segment-overlap", "segment-midoverlap",
            "annotation-overlap", "annotation-midoverlap"]
n_problems = int(sys.argv[1]) if len(sys.argv) > 1 else 60
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 7
rng = np.random.default_rng(seed0)
ctx = make_context()
cells = 0
for it in range(n_problems):
    K = int(rng.integers(1, 5))
    A = int(rng.integers(1, 40))
    S = int(rng.integers(1, 70))
    span = int(rng.choice([3000, 300000, 40000000]))
    alen = int(rng.choice([5, 90, 2000, 60000, 3000000]))
    annos = [[helpers.random_list(rng, span, int(rng.integers(0, 300)), min(alen, span // 2)) for _ in range(K)]
             for _ in range(A)]
    nseg = [int(rng.integers(0, 4)) for _ in range(K)]
    samples = [[helpers.random_list(rng, span, int(rng.integers(0, 500)), int(rng.choice([3, 40, 700, 9000, 2000000])))
                for _ in range(K)] for _ in range(S)]
    an = device.Annotations(ctx, annos, key_ws_nseg=nseg)
    got = an.count_lists(COUNTERS, samples)
    an.close()
    for s in range(S):
        exp = oracle.count_placed(samples[s], annos, nseg, COUNTERS)
        if not np.array_equal(got[:, s, :], exp):
            print("MISMATCH problem %i sample %i (K=%i A=%i span=%i alen=%i)" % (it, s, K, A, span, alen))
            bad = np.argwhere(got[:, s, :] != exp)[:5]
            print(bad, got[:, s, :][tuple(bad.T)], exp[tuple(bad.T)])
            sys.exit(1)
    cells += S * A * len(COUNTERS)
print("ok: %i problems, %i (sample, track, counter) cells identical to the oracle" % (n_problems, cells))
ctx.close()
