"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.
Bit-exact: everything here is integer / index work; nucleotide-density is a float64 sum in a fixed
order and must be bit-exact too."""
import numpy as np
import pytest

from tests import helpers

pytestmark = pytest.mark.gpu

COUNTERS = ["nucleotide-overlap", "nucleotide-density", "segment-overlap", "segment-midoverlap",
            "annotation-overlap", "annotation-midoverlap"]


def test_place_single_units_match_oracle(ctx, oracle):
    """SamplerAnnotator.sample of one unit == sequential restatement under the same Philox stream"""
    from gat_b200 import device
    rng = np.random.default_rng(11)
    ntrim = ncp = 0
    for it in range(60):
        segs, ws = helpers.random_unit(rng)
        bucket = int(rng.choice([1, 1, 1, 3, 7]))
        smp = device.Sampler(ctx, [0], 1, False, [segs], [ws], bucket_size=bucket, nbuckets=100000)
        n = 6
        placed, status = smp.place(seed=1234 + it, track=3, sample_begin=10, n_samples=n)
        smp.close()
        for s in range(n):
            exp, info = oracle.sampler_annotator_philox(segs, ws, 1234 + it, 3, 0, 10 + s, bucket_size=bucket)
            ntrim += info.ntrims
            ncp += info.ncheckpoints
            assert np.array_equal(placed[s][0], exp), (it, s, placed[s][0][:5], exp[:5])
            assert bool(status[s, 0] & 1) == (info.nunsuccessful >= 20)
    assert ntrim > 0 and ncp > 0          # the overshoot / checkpoint paths were exercised


@pytest.mark.parametrize("n_iso", [0, 3])
def test_place_problem_matches_oracle(ctx, oracle, n_iso):
    """whole samples: all units placed, isochore units merged per contig (fromIsochores)"""
    from gat_b200 import device
    rng = np.random.default_rng(5 + n_iso)
    pr = helpers.random_problem(rng, n_contigs=4, n_iso=n_iso)
    smp = device.Sampler(ctx, pr["unit_contig"], pr["n_contigs"], pr["has_isochores"],
                         pr["unit_segments"], pr["unit_workspace"])
    n = 8
    placed, status = smp.place(seed=99, track=0, sample_begin=0, n_samples=n)
    smp.close()
    for s in range(n):
        _, exp = oracle.compute_sample_philox(pr["unit_contig"], pr["unit_segments"], pr["unit_workspace"],
                                              pr["annotations"], pr["cws_nseg"], ["nucleotide-overlap"],
                                              seed=99, track=0, sample=s, has_isochores=pr["has_isochores"],
                                              return_placed=True)
        for c in range(pr["n_contigs"]):
            assert np.array_equal(placed[s][c], exp[c]), (s, c)


def test_count_lists_match_oracle(ctx, oracle):
    """all six counters on given segment sets (same-placement parity), several keys and samples"""
    from gat_b200 import device
    rng = np.random.default_rng(21)
    for it in range(6):
        K, A, S = int(rng.integers(1, 5)), int(rng.integers(1, 12)), int(rng.integers(1, 6))
        span = int(rng.choice([5000, 300000]))
        annos = [[helpers.random_list(rng, span, int(rng.integers(0, 120)), int(rng.choice([20, 2000]))) for _ in range(K)]
                 for _ in range(A)]
        nseg = [int(rng.integers(0, 4)) for _ in range(K)]
        samples = [[helpers.random_list(rng, span, int(rng.integers(0, 150)), int(rng.choice([10, 400, 5000]))) for _ in range(K)]
                   for _ in range(S)]
        an = device.Annotations(ctx, annos, key_ws_nseg=nseg)
        got = an.count_lists(COUNTERS, samples)
        an.close()
        for s in range(S):
            exp = oracle.count_placed(samples[s], annos, nseg, COUNTERS)
            assert np.array_equal(got[:, s, :], exp), (it, s, got[:, s, :], exp)


@pytest.mark.parametrize("n_iso", [0, 3])
def test_run_matches_oracle(ctx, oracle, n_iso):
    """place + count in one call == oracle computeSample for every sample; independent of batching"""
    from gat_b200 import device
    rng = np.random.default_rng(31 + n_iso)
    pr = helpers.random_problem(rng, n_contigs=3, n_iso=n_iso, n_annot=11)
    smp = device.Sampler(ctx, pr["unit_contig"], pr["n_contigs"], pr["has_isochores"],
                         pr["unit_segments"], pr["unit_workspace"])
    an = device.Annotations(ctx, pr["annotations"], key_ws_nseg=pr["cws_nseg"])
    n = 40
    res, info = smp.run(an, COUNTERS, seed=7, track=2, sample_begin=5, n_samples=n)
    ctx.set_batch_size(16)                 # 3 batches
    res2, _ = smp.run(an, COUNTERS, seed=7, track=2, sample_begin=5, n_samples=n)
    ctx.set_batch_size(0)
    for name in COUNTERS:
        assert np.array_equal(res[name], res2[name]), name
    for s in range(n):
        exp = oracle.compute_sample_philox(pr["unit_contig"], pr["unit_segments"], pr["unit_workspace"],
                                           pr["annotations"], pr["cws_nseg"], COUNTERS, seed=7, track=2,
                                           sample=5 + s, has_isochores=pr["has_isochores"])
        for i, name in enumerate(COUNTERS):
            assert np.array_equal(np.asarray(res[name][s], dtype=np.float64), exp[i]), (s, name)
    assert int(info[2]) == 0
    smp.close()
    an.close()


def test_column_stats_match_oracle(ctx, oracle):
    """expected / CI / fold / p-value bit-exact on integer columns; stddev to 1e-12 relative"""
    rng = np.random.default_rng(41)
    for l in (1, 7, 19, 20, 100, 1000, 4097):
        A = 9
        counts = rng.integers(0, 30, size=(l, A)).astype(np.uint32)
        counts[:, 0] = 5                                   # constant column
        counts[:, 1] = 0                                   # all-zero column: fold = 1
        obs = rng.integers(0, 35, size=A).astype(np.float64)
        obs[2] = counts[:, 2].max() + 10                   # above every sample
        obs[3] = -1 if False else 0.0
        got = ctx.column_stats(counts, obs, pseudo_count=1.0)
        for a in range(A):
            st = oracle.enrichment_statistics(obs[a], counts[:, a].astype(np.float64), pseudo_count=1.0)
            assert got["expected"][a] == st.expected
            assert got["lower95"][a] == st.lower95 and got["upper95"][a] == st.upper95
            assert got["fold"][a] == st.fold
            assert got["pvalue"][a] == st.pvalue, (l, a, obs[a], got["pvalue"][a], st.pvalue)
            assert got["stddev"][a] == pytest.approx(st.stddev, rel=1e-12, abs=1e-12)


def test_column_stats_float_and_reference(ctx, oracle):
    """float64 (nucleotide-density) columns and the --null fold shift"""
    rng = np.random.default_rng(43)
    l, A = 500, 6
    counts = np.round(rng.random((l, A)) * 4, 2)
    obs = np.round(rng.random(A) * 4, 2)
    obs[0] = counts[3, 0]
    got = ctx.column_stats(counts, obs)
    ref = np.array([0.5, 1.0, 2.0, 1.5, 3.0, 0.25])
    got_ref = ctx.column_stats(counts, obs, ref_fold=ref)
    for a in range(A):
        st = oracle.enrichment_statistics(obs[a], counts[:, a])
        assert got["pvalue"][a] == st.pvalue
        assert got["lower95"][a] == st.lower95 and got["upper95"][a] == st.upper95
        assert got["expected"][a] == pytest.approx(st.expected, rel=1e-13)
        assert got["stddev"][a] == pytest.approx(st.stddev, rel=1e-12)
        st = oracle.enrichment_statistics(obs[a], counts[:, a], reference_fold=ref[a])
        assert got_ref["pvalue"][a] == st.pvalue
        assert got_ref["expected"][a] == pytest.approx(st.expected, rel=1e-13)
        assert got_ref["lower95"][a] == st.lower95 and got_ref["upper95"][a] == st.upper95


def _fallback_problem(rng, oracle, big_list=False):
    """random lists; with big_list one track has > 65534 intervals on a key (no bin index possible)"""
    K, A, S = 2, 11, 3
    span = 4000000 if big_list else 300000
    annos = [[helpers.random_list(rng, span, int(rng.integers(0, 150)), 900) for _ in range(K)] for _ in range(A)]
    if big_list:
        starts = np.arange(0, 70000, dtype=np.int64) * 50
        annos[3][1] = np.stack([starts, starts + rng.integers(1, 40, len(starts))], axis=1).astype(np.uint32)
        # intervals at and beyond the 2^20 - 1 bases the packed index entry can hold
        annos[5][0] = np.array([[1000, 3500000]], dtype=np.uint32)
        annos[6][1] = np.array([[10, 10 + 1048575], [2000000, 2000000 + 1048574]], dtype=np.uint32)
    nseg = [2, 3]
    samples = [[helpers.random_list(rng, span, int(rng.integers(1, 400)), int(rng.choice([30, 700, 9000]))) for _ in range(K)]
               for _ in range(S)]
    return annos, nseg, samples


def test_count_index_geometries_match_oracle(ctx, oracle, monkeypatch):
    """every geometry of the annotation grid index gives the oracle's counts: several track groups per
    launch, bins much narrower / much wider than the intervals, an index that outgrows its estimated
    capacity and is rebuilt at the exact size, lists with > 65534 intervals, intervals longer than a packed
    entry holds, item tables of one key at a time, small CTAs"""
    from gat_b200 import device
    rng = np.random.default_rng(77)
    knobs = ("GATB_GROUP_TRACKS", "GATB_BIN_SHIFT", "GATB_INDEX_CAPACITY", "GATB_KEY_GROUP", "GATB_COUNT_THREADS",
             "GATB_SCHUNK")
    cases = [
        (False, {}),
        (False, {"GATB_GROUP_TRACKS": "4"}),
        (False, {"GATB_BIN_SHIFT": "4", "GATB_KEY_GROUP": "1"}),
        (False, {"GATB_BIN_SHIFT": "20", "GATB_COUNT_THREADS": "64"}),
        (False, {"GATB_INDEX_CAPACITY": "16", "GATB_SCHUNK": "1"}),
        (True, {}),
        (True, {"GATB_SCHUNK": "2", "GATB_BIN_SHIFT": "6"}),
        (True, {"GATB_GROUP_TRACKS": "3", "GATB_INDEX_CAPACITY": "1000", "GATB_KEY_GROUP": "1"}),
    ]
    for big_list, env in cases:
        annos, nseg, samples = _fallback_problem(rng, oracle, big_list)
        for k in knobs:
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        c2 = helpers.new_context_like(ctx)           # the knobs are read when the context / the set is created
        an = device.Annotations(c2, annos, key_ws_nseg=nseg)
        got = an.count_lists(COUNTERS, samples)
        an.close()
        c2.close()
        for s in range(len(samples)):
            exp = oracle.count_placed(samples[s], annos, nseg, COUNTERS)
            assert np.array_equal(got[:, s, :], exp), (big_list, env, s)
    for k in knobs:
        monkeypatch.delenv(k, raising=False)


def test_format_counts_matches_python_formatting(ctx):
    """gatb_format_counts: the text of the counts table, column by column, equals the reference's
    ",".join("%i" % x) (gat/__init__.py:1081-1086) -- digit boundaries, 2^32-1, one sample, chunk edges"""
    import torch
    rng = np.random.default_rng(5)
    edge = np.array([0, 9, 10, 99, 100, 999, 1000, 99999, 100000, 999999999, 1000000000, 4294967295], dtype=np.uint32)
    for S, A in ((1, 1), (1, 5), (12, 3), (256, 2), (257, 7), (1000, 4), (5000, 33)):
        m = rng.integers(0, 2 ** 32, size=(S, A), dtype=np.uint64).astype(np.uint32)
        m >>= rng.integers(0, 32, size=(S, A)).astype(np.uint32)          # every number of digits
        m[:min(S, len(edge)), 0] = edge[:min(S, len(edge))]
        want = [",".join("%i" % x for x in m[:, a]) for a in range(A)]
        text, off = ctx.format_counts(counts=m)
        got = [text[int(off[a]):int(off[a + 1])].tobytes().decode("ascii") for a in range(A)]
        assert got == want, (S, A)
        t = torch.from_numpy(m.view(np.int32)).cuda()
        text, off = ctx.format_counts(device_ptr=t.data_ptr(), n_samples=S, n_cols=A)
        got = [text[int(off[a]):int(off[a + 1])].tobytes().decode("ascii") for a in range(A)]
        assert got == want, (S, A, "device")


def test_invalid_inputs_fail_loudly(ctx):
    """error behaviour at the boundary: non-normalized lists, out-of-range coordinates, too-large segments"""
    from gat_b200 import device, _lib
    good = np.array([[10, 20], [30, 40]], dtype=np.uint32)
    with pytest.raises(_lib.GatB200Error) as e:
        device.Annotations(ctx, [[np.array([[10, 20], [15, 40]], dtype=np.uint32)]])      # overlapping
    assert e.value.code == _lib.ERR_INVALID
    with pytest.raises(_lib.GatB200Error) as e:
        device.Annotations(ctx, [[np.array([[10, 2 ** 31 + 5]], dtype=np.uint32)]])        # coordinate >= 2^31
    assert e.value.code == _lib.ERR_RANGE
    with pytest.raises(_lib.GatB200Error) as e:
        device.Sampler(ctx, [0], 1, False, [np.array([[30, 40], [10, 20]], dtype=np.uint32)], [good])   # unsorted
    assert e.value.code == _lib.ERR_INVALID
    with pytest.raises(_lib.GatB200Error) as e:
        device.Sampler(ctx, [0, 0], 1, False, [good, good], [good, good])                   # 2 units, no isochores
    assert e.value.code == _lib.ERR_INVALID
    with pytest.raises(_lib.GatB200Error) as e:
        device.Sampler(ctx, [0], 1, False, [np.array([[0, 5000]], dtype=np.uint32)],
                       [np.array([[0, 100000]], dtype=np.uint32)], bucket_size=1, nbuckets=1000)
    assert e.value.code == _lib.ERR_TOO_LARGE
    an = device.Annotations(ctx, [[good]])
    with pytest.raises(_lib.GatB200Error):
        an.count_lists(["nucleotide-density"], [[good]])          # density needs key_ws_nseg
    an.close()
    # empty annotation tracks and empty samples are fine
    an = device.Annotations(ctx, [[np.zeros((0, 2), dtype=np.uint32)], [good]], key_ws_nseg=[1])
    out = an.count_lists(COUNTERS, [[good], [np.zeros((0, 2), dtype=np.uint32)]])
    assert out[0, 0, 0] == 0 and out[0, 0, 1] == 20 and (out[:, 1, :] == 0).all()
    an.close()


def test_sampler_segments_matches_oracle(ctx, oracle):
    """SamplerSegments (exactly len(segments) placements per unit, merged per contig by fromIsochores):
    contig-level samples and counts equal the oracle's; without isochores the call is refused"""
    from gat_b200 import device, _lib
    rng = np.random.default_rng(91)
    pr = helpers.random_problem(rng, n_contigs=3, n_iso=3, n_annot=5)
    smp = device.Sampler(ctx, pr["unit_contig"], pr["n_contigs"], True, pr["unit_segments"], pr["unit_workspace"])
    smp.set_kind("segments")
    an = device.Annotations(ctx, pr["annotations"], key_ws_nseg=pr["cws_nseg"])
    n = 12
    placed, status = smp.place(seed=4, track=1, sample_begin=3, n_samples=n)
    res, _ = smp.run(an, COUNTERS, seed=4, track=1, sample_begin=3, n_samples=n)
    oracle.set_sampler_kind("segments")
    try:
        for s in range(n):
            exp, exp_placed = oracle.compute_sample_philox(pr["unit_contig"], pr["unit_segments"], pr["unit_workspace"],
                                                           pr["annotations"], pr["cws_nseg"], COUNTERS, seed=4, track=1,
                                                           sample=3 + s, has_isochores=True, return_placed=True)
            for c in range(pr["n_contigs"]):
                assert np.array_equal(placed[s][c], exp_placed[c]), (s, c)
            for i, name in enumerate(COUNTERS):
                assert np.array_equal(np.asarray(res[name][s], dtype=np.float64), exp[i]), (s, name)
    finally:
        oracle.set_sampler_kind("annotator")
    smp.close()
    an.close()
    plain = helpers.random_problem(rng, n_contigs=2, n_iso=0)
    smp = device.Sampler(ctx, plain["unit_contig"], plain["n_contigs"], False, plain["unit_segments"], plain["unit_workspace"])
    with pytest.raises(_lib.GatB200Error):
        smp.set_kind("segments")
    smp.close()


def test_sampler_shift_matches_oracle(ctx, oracle):
    """SamplerShift (gat/Engine.pyx:998-1111): single units (fragmented workspaces, several radii and
    extensions, windows without workspace) and whole problems with and without isochores equal the oracle
    under the same Philox stream -- placed segments and all counters, bit for bit"""
    from gat_b200 import device
    rng = np.random.default_rng(2027)
    npieces = noverflow = 0
    for it in range(150):
        segs, ws = helpers.random_unit(rng)
        kw = [dict(radius=2, extension=0), dict(radius=float(rng.choice([0.5, 1, 3, 7.5])), extension=0),
              dict(radius=2, extension=int(rng.choice([10, 101, 1000, 5000])))][it % 3]
        smp = device.Sampler(ctx, [0], 1, False, [segs], [ws], bucket_size=0)
        smp.set_shift(**kw)
        n = 6
        placed, status = smp.place(seed=77 + it, track=it % 4, sample_begin=50, n_samples=n)
        smp.close()
        for s in range(n):
            exp = oracle.sampler_shift(segs, ws, philox=(77 + it, it % 4, 0, 50 + s), **kw)
            if status[s, 0] & device.UNIT_OVERFLOW:
                noverflow += 1
                assert len(exp) > 8 * len(segs)         # (only a sample that really needs more room may say so)
                continue
            assert np.array_equal(placed[s][0], exp), (it, s, kw)
            npieces += len(exp)
    assert npieces > 5000 and noverflow < 20
    for n_iso, kw in ((0, dict(radius=2, extension=0)), (3, dict(radius=3, extension=0)), (2, dict(radius=2, extension=400))):
        pr = helpers.random_problem(rng, n_contigs=3, n_iso=n_iso, n_annot=5)
        has_iso = n_iso > 0
        smp = device.Sampler(ctx, pr["unit_contig"], pr["n_contigs"], has_iso, pr["unit_segments"], pr["unit_workspace"],
                             bucket_size=0)
        smp.set_shift(**kw)
        an = device.Annotations(ctx, pr["annotations"], key_ws_nseg=pr["cws_nseg"])
        n = 10
        placed, status = smp.place(seed=5, track=2, sample_begin=7, n_samples=n)
        assert not status.any()
        res, _ = smp.run(an, COUNTERS, seed=5, track=2, sample_begin=7, n_samples=n)
        oracle.set_sampler_kind("shift", **kw)
        try:
            for s in range(n):
                exp, exp_placed = oracle.compute_sample_philox(pr["unit_contig"], pr["unit_segments"], pr["unit_workspace"],
                                                               pr["annotations"], pr["cws_nseg"], COUNTERS, seed=5, track=2,
                                                               sample=7 + s, has_isochores=has_iso, return_placed=True)
                for c in range(pr["n_contigs"]):
                    assert np.array_equal(placed[s][c], exp_placed[c]), (n_iso, s, c)
                for i, name in enumerate(COUNTERS):
                    assert np.array_equal(np.asarray(res[name][s], dtype=np.float64), exp[i]), (n_iso, s, name)
        finally:
            oracle.set_sampler_kind("annotator")
        smp.close()
        an.close()


def test_async_annotations_same_counts_and_deferred_errors(ctx, oracle):
    """gatb_annotations_create_async: upload + tile build on the upload stream; the first run waits on the
    device.  Counts equal the synchronous path; invalid lists surface at wait() / at the run that used them"""
    from gat_b200 import device, _lib
    rng = np.random.default_rng(123)
    pr = helpers.random_problem(rng, n_contigs=3, n_iso=0, n_annot=9)
    smp = device.Sampler(ctx, pr["unit_contig"], pr["n_contigs"], False, pr["unit_segments"], pr["unit_workspace"])
    sync = device.Annotations(ctx, pr["annotations"], key_ws_nseg=pr["cws_nseg"])
    want, _ = smp.run(sync, COUNTERS, seed=8, track=0, sample_begin=0, n_samples=40)
    for use_wait in (False, True):
        lazy = device.Annotations(ctx, pr["annotations"], key_ws_nseg=pr["cws_nseg"], lazy=True)
        if use_wait:
            lazy.wait()
        got, _ = smp.run(lazy, COUNTERS, seed=8, track=0, sample_begin=0, n_samples=40)
        for name in COUNTERS:
            assert np.array_equal(np.asarray(got[name]), np.asarray(want[name])), (use_wait, name)
        lazy.close()
    # destroyed while still pending: must not crash or leak the validation slot
    for _ in range(300):
        device.Annotations(ctx, pr["annotations"], lazy=True).close()
    bad_lists = [[np.array([[10, 20], [15, 40]], dtype=np.uint32) for _ in range(pr["n_contigs"])]]
    bad = device.Annotations(ctx, bad_lists, lazy=True)
    with pytest.raises(_lib.GatB200Error) as e:
        bad.wait()
    assert e.value.code == _lib.ERR_INVALID
    with pytest.raises(_lib.GatB200Error):
        smp.run(bad, ["nucleotide-overlap"], seed=8, track=0, sample_begin=0, n_samples=4)
    bad.close()
    bad = device.Annotations(ctx, bad_lists, lazy=True)
    with pytest.raises(_lib.GatB200Error) as e:                      # first use reports it
        smp.run(bad, ["nucleotide-overlap"], seed=8, track=0, sample_begin=0, n_samples=4)
    assert e.value.code == _lib.ERR_INVALID
    bad.close()
    sync.close()
    smp.close()


def test_count_filter_edge_geometries(ctx, oracle):
    """the bitmap / bin-index filter at its edges: coordinates just below 2^31, an extent of a few positions,
    segments far beyond the last interval, segments longer than a bitmap window and than 32 windows, segments
    starting exactly on window boundaries, annotations touching each other across tracks"""
    from gat_b200 import device
    rng = np.random.default_rng(2718)
    top = 2 ** 31 - 1
    K = 4
    # (build [s, s+len) lists by hand so that ends stay <= 2^31 - 1)
    def near_top(n, maxlen):
        s = top - rng.integers(2, 5000000, n)
        l = rng.integers(1, maxlen + 1, n)
        e = np.minimum(s + l, top)
        return helpers.normalize(np.stack([s, e], axis=1))
    tiny = np.array([[1, 2], [3, 5]], dtype=np.uint32)
    grid = np.array([[i * 1024, i * 1024 + 1] for i in range(1, 400, 3)], dtype=np.uint32)       # window boundaries
    touching_a = np.array([[i * 100, i * 100 + 50] for i in range(200)], dtype=np.uint32)
    touching_b = np.array([[i * 100 + 50, i * 100 + 100] for i in range(200)], dtype=np.uint32)
    annos = [[near_top(300, 4000), tiny, grid, touching_a],
             [near_top(50, 100000), np.zeros((0, 2), dtype=np.uint32), helpers.random_list(rng, 500000, 300, 2000), touching_b],
             [near_top(5, 10), tiny, np.array([[0, 1]], dtype=np.uint32), helpers.random_list(rng, 20000, 100, 30)]]
    nseg = [1, 2, 3, 1]
    samples = []
    for s in range(12):
        samples.append([
            near_top(200, int(rng.choice([50, 1500, 60000]))),                                       # near 2^31
            helpers.random_list(rng, 40, 8, 6),                                                      # tiny extent
            helpers.normalize(np.stack([np.arange(0, 600) * 1024, np.arange(0, 600) * 1024 + rng.integers(1, 1025, 600)], axis=1)),
            helpers.random_list(rng, 40000, 60, int(rng.choice([20, 1025, 40000]))),                 # long segments
        ])
    samples.append([np.array([[top - 5, top]], dtype=np.uint32), np.array([[1000, 2000]], dtype=np.uint32),
                    np.array([[10 ** 9, 10 ** 9 + 5]], dtype=np.uint32), np.array([[0, 2 * 10 ** 9]], dtype=np.uint32)])
    an = device.Annotations(ctx, annos, key_ws_nseg=nseg)
    got = an.count_lists(COUNTERS, samples)
    an.close()
    for s in range(len(samples)):
        exp = oracle.count_placed(samples[s], annos, nseg, COUNTERS)
        assert np.array_equal(got[:, s, :], exp), s


def test_overlap_pieces_counter_matches_intersect(ctx, oracle):
    """GATB_OVERLAP_PIECES = len(a.intersect(b)) (with nucleotide-overlap = its sum): the overlap columns of
    AnnotatorResultExtended (gat/Engine.pyx:1911-1928) against the oracle's SegmentList.intersect"""
    from gat_b200 import device
    rng = np.random.default_rng(404)
    K, A, S = 3, 11, 7
    span = 300000
    annos = [[helpers.random_list(rng, span, int(rng.integers(0, 400)), int(rng.choice([40, 3000]))) for _ in range(K)]
             for _ in range(A)]
    samples = [[helpers.random_list(rng, span, int(rng.integers(0, 300)), int(rng.choice([25, 900, 20000]))) for _ in range(K)]
               for _ in range(S)]
    an = device.Annotations(ctx, annos)
    got = an.count_lists(["overlap-pieces", "nucleotide-overlap"], samples)
    an.close()
    for s in range(S):
        for a in range(A):
            pieces = [oracle.intersect(samples[s][k], annos[a][k]) for k in range(K)]
            assert got[0, s, a] == sum(len(p) for p in pieces), (s, a)
            assert got[1, s, a] == sum(oracle.total(p) for p in pieces), (s, a)
