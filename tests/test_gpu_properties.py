"""GPU tests at BASELINE sizes through size-independent properties (determinism, shard invariance, exact
workspace coverage, additivity over disjoint annotation splits) plus oracle spot checks on a few samples."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def big(ctx):
    """hg19-shaped: 10 000 segments x 64 annotation tracks of 20 000 intervals, 24 contigs (BASELINE config 2 shape)"""
    import gat_b200
    from gat_b200 import synthetic, device
    segments, annotations, workspaces, _ = synthetic.make(10000, 64, 20000)
    workspace = synthetic.prepare(segments, annotations, workspaces)
    problem = gat_b200.TrackProblem(segments["merged"], workspace)
    atracks, lists, nseg = gat_b200.buildContigAnnotations(annotations, workspace, problem.contigs)
    smp = device.Sampler(ctx, problem.unit_contig, len(problem.contigs), problem.has_isochores,
                         problem.unit_segments, problem.unit_workspace)
    annos = device.Annotations(ctx, lists, key_ws_nseg=nseg)
    yield dict(problem=problem, lists=lists, nseg=nseg, smp=smp, annos=annos, workspace=workspace)
    smp.close()
    annos.close()


def test_determinism_and_shard_invariance(ctx, big):
    """the matrix depends only on (seed, track, global sample index): any sharding of the sample range,
    any batch size, gives the same rows (what makes the multi-GPU result independent of the GPU count)"""
    smp, annos = big["smp"], big["annos"]
    names = ["nucleotide-overlap", "segment-overlap"]
    full, info = smp.run(annos, names, seed=11, track=0, sample_begin=0, n_samples=512)
    again, _ = smp.run(annos, names, seed=11, track=0, sample_begin=0, n_samples=512)
    lo, _ = smp.run(annos, names, seed=11, track=0, sample_begin=0, n_samples=200)
    ctx.set_batch_size(100)
    hi, _ = smp.run(annos, names, seed=11, track=0, sample_begin=200, n_samples=312)
    ctx.set_batch_size(0)
    other, _ = smp.run(annos, names, seed=12, track=0, sample_begin=0, n_samples=64)
    for n in names:
        assert np.array_equal(full[n], again[n])
        assert np.array_equal(full[n], np.concatenate([lo[n], hi[n]]))
        assert not np.array_equal(full[n][:64], other[n])
    assert int(info[2]) == 0 and int(info[0]) > 512 * 9000
    # every annotation is hit by some sample and the counts vary between samples
    assert (full["nucleotide-overlap"].max(axis=0) > 0).all()
    assert (full["nucleotide-overlap"].std(axis=0) > 0).all()


def test_placed_samples_cover_exactly_the_input(ctx, big, oracle):
    """sampler property of test/benchmark_gat.py:773-780, 828-837 at full size: every sample lies in the
    workspace, is sorted / disjoint / non-adjacent, and covers exactly the input's workspace bases"""
    pr, smp = big["problem"], big["smp"]
    placed, status = smp.place(seed=5, track=0, sample_begin=1000, n_samples=16)
    assert not (status & 2).any()
    want = [oracle.total(oracle.intersect(pr.unit_segments[u], pr.unit_workspace[u])) for u in range(len(pr.unit_contig))]
    for s in range(16):
        for u, c in enumerate(pr.unit_contig):
            a = placed[s][c].astype(np.int64)
            assert len(a) > 0
            assert (a[:, 0] < a[:, 1]).all() and (a[1:, 0] > a[:-1, 1]).all()      # merged: gaps >= 1 base
            ws = pr.unit_workspace[u].astype(np.int64)
            cov = np.clip(np.minimum(a[:, 1], ws[0, 1]) - np.maximum(a[:, 0], ws[0, 0]), 0, None)
            assert (cov > 0).all()
            if not (status[s, u] & 1):
                assert cov.sum() == want[u], (s, u)
    # two spot samples against the sequential oracle
    for s in (0, 15):
        for u, c in enumerate(pr.unit_contig):
            exp, _ = oracle.sampler_annotator_philox(pr.unit_segments[u], pr.unit_workspace[u], 5, 0, u, 1000 + s)
            assert np.array_equal(placed[s][c], exp), (s, u)


def test_counts_additive_over_disjoint_annotation_split(ctx, big):
    """split one annotation track into its even and odd intervals: nucleotide-overlap and
    annotation-overlap add up exactly, segment-overlap is sub-additive"""
    from gat_b200 import device
    pr, smp = big["problem"], big["smp"]
    base = big["lists"][0]
    even = [a[0::2] for a in base]
    odd = [a[1::2] for a in base]
    annos = device.Annotations(ctx, [base, even, odd], key_ws_nseg=big["nseg"])
    names = ["nucleotide-overlap", "annotation-overlap", "segment-overlap", "nucleotide-density"]
    res, _ = smp.run(annos, names, seed=3, track=0, sample_begin=0, n_samples=256)
    annos.close()
    for n in ("nucleotide-overlap", "annotation-overlap"):
        assert np.array_equal(res[n][:, 0], res[n][:, 1] + res[n][:, 2]), n
    so = res["segment-overlap"].astype(np.int64)
    assert (so[:, 0] <= so[:, 1] + so[:, 2]).all() and (so[:, 0] >= np.maximum(so[:, 1], so[:, 2])).all()
    # one workspace segment per contig: density == overlap
    assert np.array_equal(res["nucleotide-density"][:, 0], res["nucleotide-overlap"][:, 0].astype(np.float64))


def test_counts_match_oracle_on_spot_samples(ctx, big, oracle):
    names = ["nucleotide-overlap", "segment-overlap", "segment-midoverlap", "annotation-overlap"]
    pr, smp, annos = big["problem"], big["smp"], big["annos"]
    res, _ = smp.run(annos, names, seed=77, track=1, sample_begin=40000, n_samples=3)
    for s in range(3):
        exp = oracle.compute_sample_philox(pr.unit_contig, pr.unit_segments, pr.unit_workspace, big["lists"],
                                           big["nseg"], names, seed=77, track=1, sample=40000 + s)
        for i, n in enumerate(names):
            assert np.array_equal(res[n][s].astype(np.float64), exp[i]), (s, n)


def test_isochore_config_matches_oracle(ctx, oracle):
    """BASELINE config 3 shape: 8 GC isochores (<= 192 units), segment-overlap; spot samples vs the oracle"""
    import gat_b200
    from gat_b200 import synthetic, device
    segments, annotations, workspaces, iso = synthetic.make(10000, 8, 20000, isochores=True)
    workspace = synthetic.prepare(segments, annotations, workspaces, iso)
    pr = gat_b200.TrackProblem(segments["merged"], workspace)
    assert pr.has_isochores and 150 <= len(pr.unit_keys) <= 192 and len(pr.contigs) == 24
    atracks, lists, nseg = gat_b200.buildContigAnnotations(annotations, workspace, pr.contigs)
    smp = device.Sampler(ctx, pr.unit_contig, len(pr.contigs), True, pr.unit_segments, pr.unit_workspace)
    annos = device.Annotations(ctx, lists, key_ws_nseg=nseg)
    names = ["segment-overlap", "nucleotide-overlap", "nucleotide-density"]
    res, info = smp.run(annos, names, seed=9, track=0, sample_begin=0, n_samples=96)
    assert int(info[2]) == 0
    for s in (0, 95):
        exp = oracle.compute_sample_philox(pr.unit_contig, pr.unit_segments, pr.unit_workspace, lists, nseg, names,
                                           seed=9, track=0, sample=s, has_isochores=True)
        for i, n in enumerate(names):
            assert np.array_equal(np.asarray(res[n][s], dtype=np.float64), exp[i]), (s, n)
    smp.close()
    annos.close()


def test_column_stats_at_scale(ctx):
    """200 000 samples x 48 columns: mean exact, order statistics exact, p-value counts exact vs numpy"""
    rng = np.random.default_rng(8)
    l, A = 200000, 48
    counts = rng.poisson(rng.integers(1, 4000, A), size=(l, A)).astype(np.uint32)
    obs = np.array([np.quantile(counts[:, a], q) for a, q in zip(range(A), np.linspace(0, 1, A))]).round()
    got = ctx.column_stats(counts, obs, pseudo_count=1.0)
    srt = np.sort(counts, axis=0)
    off = int(0.05 * l)
    assert np.array_equal(got["expected"], counts.sum(axis=0, dtype=np.uint64) / l)
    assert np.array_equal(got["lower95"], srt[min(off, l - 1)].astype(np.float64))
    assert np.array_equal(got["upper95"], srt[max(l - off, 0)].astype(np.float64))
    assert np.allclose(got["stddev"], counts.std(axis=0), rtol=1e-12)
    for a in range(A):
        n_lt, n_eq = int((counts[:, a] < obs[a]).sum()), int((counts[:, a] == obs[a]).sum())
        if n_lt == l:
            k = 1
        elif obs[a] > got["expected"][a]:
            k = l - n_lt if (n_eq > 0 and n_lt > 0) else l - n_lt - 1
        else:
            k = n_lt + n_eq
        assert got["pvalue"][a] == max(1.0 / l, k / l), a


def test_large_units_match_oracle(ctx, oracle):
    """tutorial-scale lists (HepG2 DHS x Jurkat DHS shape: ~150 000 segments, ~160 000-interval tracks): units of
    ~12 000 segments (32 768-slot buffers, multi-stage sorts), filters with > 10 000 union intervals"""
    import gat_b200
    from gat_b200 import synthetic, device
    segments, annotations, workspaces, _ = synthetic.make(150000, 2, 160000)
    workspace = synthetic.prepare(segments, annotations, workspaces)
    pr = gat_b200.TrackProblem(segments["merged"], workspace)
    atracks, lists, nseg = gat_b200.buildContigAnnotations(annotations, workspace, pr.contigs)
    assert max(len(x) for x in pr.unit_segments) > 10000
    smp = device.Sampler(ctx, pr.unit_contig, len(pr.contigs), False, pr.unit_segments, pr.unit_workspace)
    annos = device.Annotations(ctx, lists, key_ws_nseg=nseg)
    names = ["nucleotide-overlap", "segment-overlap", "annotation-overlap"]
    res, info = smp.run(annos, names, seed=2, track=0, sample_begin=0, n_samples=24)
    assert int(info[2]) == 0
    exp = oracle.compute_sample_philox(pr.unit_contig, pr.unit_segments, pr.unit_workspace, lists, nseg, names,
                                       seed=2, track=0, sample=23)
    for i, n in enumerate(names):
        assert np.array_equal(res[n][23].astype(np.float64), exp[i]), n
    smp.close()
    annos.close()
