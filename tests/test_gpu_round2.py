"""GPU tests of the round-2 additions: buffers that grow on demand (skewed units), SamplerShift without a length
histogram, unit-level placement (per-key samples, SamplerSegments.sample), getEmpiricalPValue against the stored
expectation, the per-key --output-samples-pattern file."""
import numpy as np
import pytest

from tests import helpers

pytestmark = pytest.mark.gpu


def test_skewed_unit_grows_its_buffer(ctx, oracle):
    """HistogramSampler never draws the largest rank, so a unit whose largest segment holds most of
    the bases is refilled with the small lengths: 200 + 300 + 50 000 bp need > 150 placements against the 128
    slots a 3-segment unit used to get.  The reference grows its lists on demand; so must the library."""
    from gat_b200 import device
    ws = np.array([[0, 4000000]], dtype=np.uint32)
    segs = np.array([[1000, 1200], [5000, 5300], [100000, 150000]], dtype=np.uint32)
    smp = device.Sampler(ctx, [0], 1, False, [segs], [ws])
    n = 5
    placed, status = smp.place(seed=8, track=0, sample_begin=0, n_samples=n)
    assert not (status & device.UNIT_OVERFLOW).any()
    nmax = 0
    for s in range(n):
        exp, info = oracle.sampler_annotator_philox(segs, ws, 8, 0, 0, s, cap=8192)
        assert np.array_equal(placed[s][0], exp), s
        nmax = max(nmax, info.nplaced)
    assert nmax > 128
    # forced: a capacity estimate that is far too small is repaired by the grow-and-repeat path of gatb_run
    # and of gatb_sampler_place (same results, larger capacity afterwards)
    rng = np.random.default_rng(3)
    pr = helpers.random_problem(rng, n_contigs=2, n_iso=0, nseg=300)
    annos = device.Annotations(ctx, pr["annotations"], key_ws_nseg=pr["cws_nseg"])
    for c in range(pr["n_contigs"]):        # skew every unit: one huge segment next to the short ones
        u = pr["unit_segments"][c]
        last = int(u[-1, 1])
        pr["unit_segments"][c] = np.concatenate([u, np.array([[last + 10, last + 60000]], dtype=np.uint32)])
        w = pr["unit_workspace"][c]
        pr["unit_workspace"][c] = helpers.normalize(np.concatenate([w, np.array([[last, last + 70000]], dtype=np.uint32)]))
    smp2 = device.Sampler(ctx, pr["unit_contig"], pr["n_contigs"], False, pr["unit_segments"], pr["unit_workspace"])
    names = ["nucleotide-overlap", "segment-overlap"]
    res, info = smp2.run(annos, names, seed=2, track=0, sample_begin=0, n_samples=16)
    assert int(info[2]) == 0
    for s in (0, 15):
        exp = oracle.compute_sample_philox(pr["unit_contig"], pr["unit_segments"], pr["unit_workspace"],
                                           pr["annotations"], pr["cws_nseg"], names, seed=2, track=0, sample=s,
                                           cap=1 << 16)
        for i, name in enumerate(names):
            assert np.array_equal(res[name][s].astype(np.float64), exp[i]), (s, name)
    smp.close()
    smp2.close()
    annos.close()


def test_overflow_growth_is_exercised(ctx, oracle, monkeypatch):
    """the grow-and-repeat path itself: GATB_PLACE_CAP_MAX clamps the initial capacities so that units DO
    overflow on the first pass; results must equal an unclamped run and the oracle"""
    from gat_b200 import device
    rng = np.random.default_rng(17)
    pr = helpers.random_problem(rng, n_contigs=3, n_iso=3, nseg=200, n_annot=3)
    annos = device.Annotations(ctx, pr["annotations"], key_ws_nseg=pr["cws_nseg"])
    names = ["nucleotide-overlap"]
    plain = device.Sampler(ctx, pr["unit_contig"], pr["n_contigs"], True, pr["unit_segments"], pr["unit_workspace"])
    want, _ = plain.run(annos, names, seed=6, track=0, sample_begin=0, n_samples=40)
    want_placed, _ = plain.place(seed=6, track=0, sample_begin=0, n_samples=4)
    cap0 = plain.capacity
    plain.close()
    monkeypatch.setenv("GATB_PLACE_CAP_MAX", "64")
    small = device.Sampler(ctx, pr["unit_contig"], pr["n_contigs"], True, pr["unit_segments"], pr["unit_workspace"])
    monkeypatch.delenv("GATB_PLACE_CAP_MAX")
    assert small.capacity < cap0
    got_placed, status = small.place(seed=6, track=0, sample_begin=0, n_samples=4)
    assert small.capacity > 64 * 3 and not (status & device.UNIT_OVERFLOW).any()
    for s in range(4):
        for c in range(pr["n_contigs"]):
            assert np.array_equal(got_placed[s][c], want_placed[s][c])
    small.close()
    monkeypatch.setenv("GATB_PLACE_CAP_MAX", "64")
    small = device.Sampler(ctx, pr["unit_contig"], pr["n_contigs"], True, pr["unit_segments"], pr["unit_workspace"])
    monkeypatch.delenv("GATB_PLACE_CAP_MAX")
    got, info = small.run(annos, names, seed=6, track=0, sample_begin=0, n_samples=40)
    assert np.array_equal(got[names[0]], want[names[0]]) and int(info[2]) == 0
    units, _ = small.place_units(seed=6, track=0, sample_begin=0, n_samples=2)
    for u in range(len(pr["unit_contig"])):
        exp, _ = oracle.sampler_annotator_philox(pr["unit_segments"][u], pr["unit_workspace"][u], 6, 0, u, 1)
        assert np.array_equal(units[1][u], exp), u
    small.close()
    annos.close()


def test_shift_sampler_needs_no_histogram(ctx, oracle):
    """SamplerShift never calls getLengthDistribution, so a 100 000 bp segment (automatic bucket
    size -> index == nbuckets -> "segment too large" for the annotator) must not stop --sampler=shift"""
    import gat_b200
    from gat_b200 import device, _lib, engine
    from gat_b200.segmentlist import SegmentList
    ws = np.array([[0, 3000000]], dtype=np.uint32)
    segs = np.array([[10000, 110000], [500000, 500400], [900000, 1000000]], dtype=np.uint32)
    with pytest.raises(_lib.GatB200Error) as e:
        device.Sampler(ctx, [0], 1, False, [segs], [ws], bucket_size=0, nbuckets=100000)
    assert e.value.code == _lib.ERR_TOO_LARGE
    smp = device.Sampler(ctx, [0], 1, False, [segs], [ws], bucket_size=1, nbuckets=0)
    with pytest.raises(_lib.GatB200Error):          # no histogram: the annotator cannot run on it
        smp.place(seed=1, track=0, sample_begin=0, n_samples=1)
    smp.set_shift(2, 0)
    placed, status = smp.place(seed=1, track=0, sample_begin=0, n_samples=4)
    smp.close()
    for s in range(4):
        assert np.array_equal(placed[s][0], oracle.sampler_shift(segs, ws, radius=2, philox=(1, 0, 0, s)))
    engine.seed(5)
    sl, w = SegmentList(array=segs), SegmentList(array=ws)
    sl._normalized = w._normalized = True
    out = gat_b200.SamplerShift(radius=2).sample(sl, w)
    assert out.sum() == 200400


def test_unit_level_placement_and_sampler_segments_sample(ctx, oracle):
    """gatb_sampler_place_units returns what sampler.sample(segs[key], workspace[key]) returns for every key
    (gat/__init__.py:531-546): annotator units sorted + merged, SamplerSegments units in draw order"""
    import gat_b200
    from gat_b200 import device, engine
    from gat_b200.segmentlist import SegmentList
    rng = np.random.default_rng(12)
    for n_iso in (0, 3):
        pr = helpers.random_problem(rng, n_contigs=3, n_iso=n_iso)
        smp = device.Sampler(ctx, pr["unit_contig"], pr["n_contigs"], pr["has_isochores"],
                             pr["unit_segments"], pr["unit_workspace"])
        units, _ = smp.place_units(seed=21, track=2, sample_begin=5, n_samples=3)
        for s in range(3):
            assert len(units[s]) == len(pr["unit_contig"])
            for u in range(len(pr["unit_contig"])):
                exp, _ = oracle.sampler_annotator_philox(pr["unit_segments"][u], pr["unit_workspace"][u], 21, 2, u, 5 + s)
                assert np.array_equal(units[s][u], exp), (n_iso, s, u)
        if n_iso:
            smp.set_kind("segments")
            units, _ = smp.place_units(seed=21, track=2, sample_begin=5, n_samples=2)
            for u in range(len(pr["unit_contig"])):
                exp = oracle.sampler_segments(pr["unit_segments"][u], pr["unit_workspace"][u], philox=(21, 2, u, 6))
                assert np.array_equal(units[1][u], exp), u
        smp.close()
    # the per-call protocol of the reference: SamplerSegments().sample(segments, workspace)
    segs, ws = helpers.random_unit(rng)
    engine.seed(99)
    sl, w = SegmentList(array=segs), SegmentList(array=ws)
    sl._normalized = w._normalized = True
    sampler = gat_b200.SamplerSegments()
    first = sampler.sample(sl, w)
    second = sampler.sample(sl, w)
    assert np.array_equal(first.asarray(), oracle.sampler_segments(segs, ws, philox=(99, 0xFFFFFF, 0, 0)))
    assert np.array_equal(second.asarray(), oracle.sampler_segments(segs, ws, philox=(99, 0xFFFFFF, 0, 1)))


def test_empirical_pvalue_uses_stored_expectation(ctx, oracle):
    """with a --null reference the stored expectation is mean * fold; getEmpiricalPValue must pick the
    over- / under-representation branch against it (gat/Engine.pyx:1564), not against a recomputed mean"""
    from gat_b200 import engine

    class Ref(object):
        fold = 3.0

    rng = np.random.default_rng(4)
    samples = rng.poisson(100, 400).astype(np.float64)
    r = engine.AnnotatorResult("t", "a", "nucleotide-overlap", 150.0, samples, reference=Ref())
    assert r.expected == pytest.approx(samples.mean() * 3.0)
    srt = np.sort(samples)
    for value in (90.0, 120.0, 200.0, 310.0, float(srt[10]), float(srt[-3])):
        # getTwoSidedPValue restated on the sorted samples with the stored expectation
        l = len(srt)
        idx = int(np.searchsorted(srt, value, side="left"))
        if idx == l:
            idx = 1
        elif value > r.expected:
            while idx > 0 and srt[idx] == value:
                idx -= 1
            idx = l - (idx + 1)
        else:
            while idx < l and srt[idx] == value:
                idx += 1
        assert r.getEmpiricalPValue(value) == max(1.0 / l, idx / l), value
    # without a reference it equals the statistic computed at construction
    r2 = engine.AnnotatorResult("t", "a", "nucleotide-overlap", 93.0, samples)
    assert r2.getEmpiricalPValue(93.0) == r2.pvalue == oracle.enrichment_statistics(93.0, samples).pvalue


def test_output_samples_pattern_has_unit_keys(ctx, oracle, tmp_path):
    """--output-samples-pattern writes every unit under its own key (contig.isochore), as the reference does
    (gat/__init__.py:518-559); the blocks are the oracle's per-unit placements"""
    import gat_b200
    from gat_b200 import synthetic, engine
    genome = [("chrA", 400000), ("chrB", 250000)]
    segments, annotations, workspaces, iso = synthetic.make(300, 3, 200, isochores=True, genome=genome,
                                                            isochore_tile=20000, n_isochores=3)
    workspace = synthetic.prepare(segments, annotations, workspaces, iso)
    engine.seed(31)
    pattern = str(tmp_path / "samples_%s.bed")
    gat_b200.run(segments, annotations, workspace, gat_b200.SamplerAnnotator(), [engine.CounterNucleotideOverlap()],
                 engine.UnconditionalWorkspace(), num_samples=3, output_samples_pattern=pattern)
    pr = gat_b200.TrackProblem(segments["merged"], workspace)
    blocks, cur = {}, None
    with open(str(tmp_path / "samples_merged.bed")) as f:
        for line in f:
            if line.startswith("track name="):
                cur = int(line.strip().split("=")[1])
                blocks[cur] = {}
            else:
                key, s, e = line.rstrip("\n").split("\t")
                blocks[cur].setdefault(key, []).append((int(s), int(e)))
    assert sorted(blocks) == [0, 1, 2]
    assert any("." in k for k in blocks[0])
    for s in range(3):
        for u, key in enumerate(pr.unit_keys):
            exp, _ = oracle.sampler_annotator_philox(pr.unit_segments[u], pr.unit_workspace[u], 31, 0, u, s)
            got = np.array(blocks[s].get(key, []), dtype=np.uint32).reshape(-1, 2)
            assert np.array_equal(got, exp), (s, key)


def test_multirank_run_equals_single_rank():
    """gat_b200.run under torch.distributed.run over NCCL (2 ranks, S not divisible by 2): statistics and sample
    matrices equal the 1-rank run for the all-gather and the column-sharded route (tools/multirank_check.py).
    Needs two GPUs: skipped on a one-GPU box, run explicitly with `gpurun --gpus 2` (profiles/r02_multirank_*.json)."""
    import json
    import os
    import subprocess
    import sys
    import torch
    from tests.test_parallel_gloo import free_port
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(free_port()),
                          os.path.join(root, "tools", "multirank_check.py"), "--samples", "2001", "--tracks", "12"],
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout + out.stderr
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    assert json.loads(line)["ok"]


def test_output_routes_deliver_rows_and_column_blocks(ctx, oracle):
    """gatb_set_output_routes: the counting kernel's epilogue writes every finished row to several destinations --
    whole rows at a row offset (the all-gather layout) and column blocks with their own row stride (the all-to-all
    by column layout), several counter planes, several internal batches; same numbers as the plain output"""
    import torch
    from gat_b200 import device, _lib
    rng = np.random.default_rng(77)
    pr = helpers.random_problem(rng, n_contigs=3, n_iso=0, n_annot=11)
    A, S = 11, 50
    smp = device.Sampler(ctx, pr["unit_contig"], pr["n_contigs"], False, pr["unit_segments"], pr["unit_workspace"])
    annos = device.Annotations(ctx, pr["annotations"], key_ws_nseg=pr["cws_nseg"])
    names = ["nucleotide-overlap", "segment-overlap"]
    want, _ = smp.run(annos, names, seed=9, track=0, sample_begin=100, n_samples=S)
    dev = torch.device("cuda", 0)
    full = torch.full((2, S + 7, A), -1, dtype=torch.int32, device=dev)         # rows 5 .. 5 + S of a taller matrix
    left = torch.full((2, S, 4), -1, dtype=torch.int32, device=dev)             # columns 0..3
    right = torch.full((2, S, 7), -1, dtype=torch.int32, device=dev)            # columns 4..10
    routes = [dict(base=full.data_ptr(), plane_stride=(S + 7) * A, row_stride=A, row0=5, col_begin=0, col_end=A),
              dict(base=left.data_ptr(), plane_stride=S * 4, row_stride=4, row0=0, col_begin=0, col_end=4),
              dict(base=right.data_ptr(), plane_stride=S * 7, row_stride=7, row0=0, col_begin=4, col_end=A)]
    for by_kernel, batch in ((False, 16), (True, 16), (False, 0)):      # copy engines / kernel stores; 4 batches / 1
        for t in (full, left, right):
            t.fill_(-1)
        ctx.set_batch_size(batch)
        ctx.set_route_mode(by_kernel)
        ctx.set_output_routes(routes)
        try:
            smp.run(annos, names, seed=9, track=0, sample_begin=100, n_samples=S, out_counts_ptr=full.data_ptr())
            for t in (left, right):                         # host outputs AND routes at once
                t.fill_(-1)
            both, _ = smp.run(annos, names, seed=9, track=0, sample_begin=100, n_samples=S)
            assert all(np.array_equal(both[n], want[n]) for n in names)
            with pytest.raises(_lib.GatB200Error):          # integer counters only
                smp.run(annos, ["nucleotide-density"], seed=9, track=0, sample_begin=100, n_samples=S)
        finally:
            ctx.set_output_routes([])
            ctx.set_route_mode(False)
            ctx.set_batch_size(0)
        torch.cuda.synchronize()
        for i, n in enumerate(names):
            w = want[n].astype(np.int64)
            assert np.array_equal(full[i, 5:5 + S].cpu().numpy(), w), (by_kernel, n)
            assert (full[i, :5] == -1).all() and (full[i, 5 + S:] == -1).all()
            assert np.array_equal(left[i].cpu().numpy(), w[:, :4]) and np.array_equal(right[i].cpu().numpy(), w[:, 4:]), n
    again, _ = smp.run(annos, names, seed=9, track=0, sample_begin=100, n_samples=S)    # routes cleared: plain output
    assert np.array_equal(again[names[0]], want[names[0]])
    smp.close()
    annos.close()


def test_streamed_column_stats_exact_and_layout_independent(ctx):
    """K5 as TMA-streamed passes (stats_stream.cu): sums, p-value counts and order statistics exact against numpy for
    shapes around every boundary of the kernel (1 column ... more columns than threads, one partial chunk ... many
    super-chunks, values up to 2^32 - 1), stddev to 1e-12, and every column's statistics identical bit for bit
    whether the column is part of the full matrix or of a column block (the column-sharded layout of run())"""
    rng = np.random.default_rng(77)
    shapes = [(1, 1), (5, 3), (1000, 50), (4097, 9), (5000, 125), (3000, 257), (2500, 1000), (1500, 1200), (70000, 16)]
    for l, A in shapes:
        hi = int(rng.choice([3, 40, 70000, 2 ** 32 - 1]))
        counts = rng.integers(0, hi, size=(l, A), endpoint=True).astype(np.uint32)
        if A > 2:
            counts[:, 1] = 7                                       # constant column
            counts[:, 2] = 0
        obs = np.array([float(np.quantile(counts[:, a], q)) for a, q in zip(range(A), np.linspace(0, 1, A))]).round()
        got = ctx.column_stats(counts, obs, pseudo_count=1.0)
        srt = np.sort(counts, axis=0)
        off = int(0.05 * l)
        lo, hi_ = (min(off, l - 1), l - off) if off > 0 else (0, l - 1)
        assert np.array_equal(got["expected"], counts.sum(axis=0, dtype=np.uint64).astype(np.float64) / l), (l, A)
        assert np.array_equal(got["lower95"], srt[lo].astype(np.float64)), (l, A)
        assert np.array_equal(got["upper95"], srt[hi_].astype(np.float64)), (l, A)
        assert np.allclose(got["stddev"], counts.astype(np.float64).std(axis=0), rtol=1e-12, atol=0), (l, A)
        for a in range(A):
            n_lt, n_eq = int((counts[:, a] < obs[a]).sum()), int((counts[:, a] == obs[a]).sum())
            if n_lt == l:
                k = 1
            elif obs[a] > got["expected"][a]:
                k = l - n_lt if (n_eq > 0 and n_lt > 0) else l - n_lt - 1
            else:
                k = n_lt + n_eq
            assert got["pvalue"][a] == max(1.0 / l, k / l), (l, A, a)
        if A >= 8:
            b0, b1 = A // 3, A // 3 + max(1, A // 4)
            block = ctx.column_stats(np.ascontiguousarray(counts[:, b0:b1]), obs[b0:b1], pseudo_count=1.0)
            for key in ("expected", "stddev", "lower95", "upper95", "fold", "pvalue"):
                assert np.array_equal(block[key], got[key][b0:b1]), (l, A, key)


def test_column_stats_do_not_depend_on_the_kernel_that_computes_them(ctx, monkeypatch):
    """a uint32 matrix is reduced by the TMA-streamed passes, by the same passes without TMA when its base is not
    16-byte aligned (a plane of a [counter][sample][column] tensor), or by the column-tiled kernels (very wide
    matrices, GATB_STATS_STREAM=0): every output must agree bit for bit, stddev included -- a 2-GPU run found the
    column-sharded statistics differing from the 1-GPU ones in the last bit of stddev when the paths did not"""
    emulated = hasattr(ctx.lib, "gatb_emulation_marker")
    monkeypatch.setenv("GATB_STATS_STREAM", "0")
    tiled = helpers.new_context_like(ctx)
    monkeypatch.delenv("GATB_STATS_STREAM")
    rng = np.random.default_rng(19)
    try:
        for l, A in ((3001, 10), (3001, 20), (1000, 33), (4100, 1000)):
            scale = rng.choice([1, 5, 300, 70000, 4000000, 2 ** 32 - 1], size=A)
            vals = (rng.random((l, A)) * scale).astype(np.uint32)
            obs = vals[0].astype(np.float64)
            a = ctx.column_stats(vals, obs, pseudo_count=1.0)
            c = tiled.column_stats(vals, obs, pseudo_count=1.0)
            if emulated:                                   # (emulated device memory is host memory)
                hold = np.zeros(l * A + 4, dtype=np.uint32)
                hold[1:1 + l * A] = vals.ravel()
                ptr = hold.ctypes.data + 4
            else:
                import torch
                hold = torch.zeros(l * A + 4, dtype=torch.int32, device="cuda")
                hold[1:1 + l * A] = torch.from_numpy(vals.view(np.int32).ravel()).cuda()
                torch.cuda.synchronize()
                ptr = hold.data_ptr() + 4
            assert ptr % 16 == 4
            b = ctx.column_stats(None, obs, device_ptr=ptr, n_samples=l, n_cols=A, is_float=False)
            for key in ("expected", "stddev", "lower95", "upper95", "fold", "pvalue"):
                assert np.array_equal(a[key], b[key]), (l, A, key, "unaligned")
                assert np.array_equal(a[key], c[key]), (l, A, key, "column-tiled")
    finally:
        tiled.close()
