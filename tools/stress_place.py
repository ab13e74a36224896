"""Randomised placement parity beyond the test suite's budget: many random units (tests/helpers.random_unit),
every sample compared bit for bit with the oracle under the same Philox stream -- verification aid

    python tools/stress_place.py [n_units] [seed]
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

if "--shift" in sys.argv:          # SamplerShift instead: python tools/stress_place.py --shift [n_units] [seed]
    sys.argv.remove("--shift")
    n_units = int(sys.argv[1]) if len(sys.argv) > 1 else 600
    seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 2024
    rng = np.random.default_rng(seed0)
    ctx = make_context()
    npieces = nover = 0
    for it in range(n_units):
        segs, ws = helpers.random_unit(rng)
        kw = [dict(radius=2, extension=0), dict(radius=float(rng.choice([0.25, 0.5, 1, 3, 7.5, 40])), extension=0),
              dict(radius=2, extension=int(rng.choice([1, 10, 101, 1000, 5000, 200000])))][it % 3]
        smp = device.Sampler(ctx, [0], 1, False, [segs], [ws], bucket_size=0)
        smp.set_shift(**kw)
        placed, status = smp.place(seed=seed0 + it, track=it % 5, sample_begin=100, n_samples=8)
        smp.close()
        for s in range(8):
            exp = oracle.sampler_shift(segs, ws, philox=(seed0 + it, it % 5, 0, 100 + s), **kw)
            if status[s, 0] & device.UNIT_OVERFLOW:
                nover += 1
                continue
            npieces += len(exp)
            if not np.array_equal(placed[s][0], exp):
                print("MISMATCH unit %i sample %i" % (it, s), kw, placed[s][0][:4], exp[:4])
                sys.exit(1)
    print("ok: SamplerShift, %i units x 8 samples identical to the oracle (%i pieces, %i samples over capacity)"
          % (n_units, npieces, nover))
    ctx.close()
    sys.exit(0)

def fragmented_unit(rng):
    """a unit whose workspace has many pieces (8 .. 600: clustered, evenly tiled or with huge gaps) -- the shapes that
    search the pieces through the bucket tables (common.cuh WS_NB)"""
    from gat_b200.segmentlist import SegmentList
    while True:
        npieces = int(rng.choice([8, 9, 31, 64, 160, 257, 600]))
        kind = int(rng.integers(0, 3))
        if kind == 0:                       # tiles with random gaps
            gaps = rng.integers(1, 30000, npieces)
            lens = rng.integers(1, 20000, npieces)
        elif kind == 1:                     # two far-apart clusters of small pieces
            gaps = rng.integers(1, 300, npieces)
            gaps[npieces // 2] = int(rng.integers(10 ** 6, 10 ** 8))
            lens = rng.integers(200, 3000, npieces)
        else:                               # a few huge pieces among tiny ones
            gaps = rng.integers(1, 2000, npieces)
            lens = rng.integers(1, 50, npieces)
            lens[rng.integers(0, npieces, 3)] = rng.integers(10 ** 5, 10 ** 7, 3)
        start = np.cumsum(gaps + np.concatenate([[0], lens[:-1]]))
        ws = np.stack([start, start + lens], axis=1).astype(np.uint32)
        span = int(ws[-1, 1])
        if span >= 2 ** 31 - 10 ** 5:
            continue
        n = int(rng.integers(1, 120))
        pos = rng.integers(0, span, n)
        ln = rng.integers(1, int(rng.choice([50, 500, 5000])) + 1, n)
        segs = helpers.normalize(np.stack([pos, pos + ln], axis=1))
        t = SegmentList(array=segs); t._normalized = True
        w = SegmentList(array=ws); w._normalized = True
        t.filter(w)
        if len(t):
            return segs, ws


fragmented = "--fragmented" in sys.argv      # python tools/stress_place.py --fragmented [n_units] [seed]
if fragmented:
    sys.argv.remove("--fragmented")
n_units = int(sys.argv[1]) if len(sys.argv) > 1 else 600
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 2024
rng = np.random.default_rng(seed0)
ctx = make_context()
ntrim = ncp = nround = 0
for it in range(n_units):
    segs, ws = fragmented_unit(rng) if fragmented else helpers.random_unit(rng)
    bucket = int(rng.choice([1, 1, 1, 3, 7]))
    smp = device.Sampler(ctx, [0], 1, False, [segs], [ws], bucket_size=bucket, nbuckets=100000)
    n = 8
    placed, status = smp.place(seed=seed0 + it, track=it % 5, sample_begin=100, n_samples=n)
    smp.close()
    for s in range(n):
        exp, info = oracle.sampler_annotator_philox(segs, ws, seed0 + it, it % 5, 0, 100 + s, bucket_size=bucket)
        ntrim += info.ntrims
        ncp += info.ncheckpoints
        nround += info.nunsuccessful >= 20
        if not np.array_equal(placed[s][0], exp):
            print("MISMATCH unit %i sample %i" % (it, s), placed[s][0][:4], exp[:4])
            sys.exit(1)
print("ok: %i %sunits x 8 samples identical to the oracle (%i trims, %i checkpoints, %i round caps)"
      % (n_units, "fragmented-workspace " if fragmented else "", ntrim, ncp, nround))
ctx.close()
