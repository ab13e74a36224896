"""GPU tests against the reference's golden vectors (tests/golden/, produced by the compiled reference):
observed counts bit-exact, same-placement per-sample counts bit-exact, sampled distributions
statistically equivalent (KS, 3 SE, binomial CI), and the gat.run()-style API end to end."""
import io
import os

import numpy as np
import pytest

from tests import golden_util as G

pytestmark = pytest.mark.gpu

COUNTERS = ["nucleotide-overlap", "nucleotide-density", "segment-overlap", "segment-midoverlap",
            "annotation-overlap", "annotation-midoverlap"]


def counters_of(names):
    from gat_b200 import engine as Engine
    return [Engine.COUNTER_CLASSES[n]() for n in names]


def load_prepared(name):
    z, meta = G.load_npz(name)
    segments = G.collection(z, "segments", meta["segments"], coded=True)
    annotations = G.collection(z, "annotations", meta["annotations"], coded=True)
    workspace = G.dictionary(z, "workspace", meta["workspace"], coded=True)
    return segments, annotations, workspace, meta


def test_observed_counts_reference_golden_run(ctx):
    """the 28 observed values of the reference's golden run (test/data/output_single.tsv; the check of
    test/check_run.py:108-112) through Engine.computeCounts on the GPU"""
    from gat_b200 import engine as Engine
    segments, annotations, workspace, meta = load_prepared("observed_testdata")
    counts = Engine.computeCounts(Engine.CounterNucleotideOverlap(), sum, segments, annotations, workspace,
                                  Engine.UnconditionalWorkspace())
    got = dict(("%s|%s" % (t, a), v) for t, r in counts.items() for a, v in r.items())
    assert len(got) == 28 and got == meta["golden"]
    assert all(isinstance(v, int) for v in got.values())


def test_observed_count_tutorial(ctx):
    """BASELINE config 1: SRF x Jurkat DHS, observed 20183 (doc/tutorialIntervalOverlap.rst:103)"""
    from gat_b200 import engine as Engine
    segments, annotations, workspace, meta = load_prepared("observed_tutorial")
    counts = Engine.computeCounts(Engine.CounterNucleotideOverlap(), sum, segments, annotations, workspace,
                                  Engine.UnconditionalWorkspace())
    (track, r), = counts.items()
    assert list(r.values()) == [20183]


def test_tutorial_run_matches_published_statistics(ctx):
    """config 1 end to end, 1000 samples: observed exact; expected / stddev / fold agree with the published
    table (246.565, 105.59, 81.53) within Monte-Carlo error; p = 1/1000"""
    import gat_b200
    from gat_b200 import engine as Engine
    segments, annotations, workspace, meta = load_prepared("observed_tutorial")
    Engine.seed(1)
    res = gat_b200.run(segments, annotations, workspace, Engine.SamplerAnnotator(bucket_size=1, nbuckets=100000),
                       [Engine.CounterNucleotideOverlap()], Engine.UnconditionalWorkspace(), num_samples=1000)
    assert len(res) == 1
    r, pub = res[0], meta["published"]
    assert r.observed == 20183 and r.nsamples == 1000
    se = pub["stddev"] / np.sqrt(1000) * np.sqrt(2)          # two independent 1000-sample estimates
    assert abs(r.expected - pub["expected"]) < 4 * se
    assert abs(r.stddev - pub["stddev"]) / pub["stddev"] < 0.15
    assert abs(r.fold - pub["fold"]) / pub["fold"] < 0.08
    assert r.pvalue == 1e-3
    row = str(r).split("\t")
    assert row[2] == "20183" and row[9] == "1.0000e-03" and len(row) == len(Engine.AnnotatorResultExtended.headers)


def _track_problem(z, tag, m):
    import gat_b200
    segments = G.collection(z, tag + "/segments", m["segments"])
    annotations = G.collection(z, tag + "/annotations", m["annotations"])
    workspace = G.dictionary(z, tag + "/workspace", m["workspace"])
    problem = gat_b200.TrackProblem(segments["merged"], workspace)
    atracks, lists, nseg = gat_b200.buildContigAnnotations(annotations, workspace, problem.contigs)
    return segments, annotations, workspace, problem, atracks, lists, nseg


@pytest.mark.parametrize("tag", ["plain", "iso"])
def test_same_placement_counts_match_reference(ctx, tag):
    """feed the reference's OWN placed samples (its --output-samples-pattern dump) to the counting kernel:
    per-sample counts of all six counters must equal the reference's, bit for bit"""
    from gat_b200 import device
    from gat_b200.segmentlist import SegmentList
    z, meta = G.load_npz("run_small")
    m = meta[tag]
    _, _, _, problem, atracks, lists, nseg = _track_problem(z, tag, m)
    samples = []
    for placed in m["placed"]:
        per = dict((c, []) for c in problem.contigs)
        for key, segs in placed.items():
            per[key.split(".")[0]].extend(segs)
        row = []
        for c in problem.contigs:
            s = SegmentList(array=np.array(per[c], dtype=np.uint32).reshape(-1, 2))
            if tag == "iso":
                s.merge(0)                     # sample.fromIsochores() (gat/__init__.py:563)
            row.append(s.asarray())
        samples.append(row)
    annos = device.Annotations(ctx, lists, key_ws_nseg=nseg)
    got = annos.count_lists(COUNTERS, samples)
    annos.close()
    for ci, counter in enumerate(COUNTERS):
        for ai, anno in enumerate(atracks):
            want = np.array(m["results"][counter][anno]["samples"])
            assert np.array_equal(got[ci, :, ai], want), (counter, anno)


@pytest.mark.parametrize("tag", ["plain", "iso"])
def test_observed_counts_all_counters_match_reference(ctx, tag):
    """observed counts of all six counters incl. the isochore convention (per-isochore keys, segments
    filtered not truncated -- quirk C-7) equal the reference's"""
    from gat_b200 import engine as Engine
    z, meta = G.load_npz("run_small")
    m = meta[tag]
    segments, annotations, workspace, _, atracks, _, _ = _track_problem(z, tag, m)
    for counter in counters_of(COUNTERS):
        counts = Engine.computeCounts(counter, sum, segments, annotations, workspace, Engine.UnconditionalWorkspace())
        for anno in atracks:
            assert counts["merged"][anno] == m["results"][counter.name][anno]["observed"], (counter.name, anno)


def ks_statistic(a, b):
    a, b = np.sort(a), np.sort(b)
    allv = np.concatenate([a, b])
    ca = np.searchsorted(a, allv, side="right") / len(a)
    cb = np.searchsorted(b, allv, side="right") / len(b)
    return np.abs(ca - cb).max()


@pytest.mark.parametrize("tag", ["plain", "iso"])
def test_sampled_distributions_equivalent_to_reference(ctx, tag):
    """stochastic outputs differ only through the RNG: per annotation, two-sample KS against 2000
    reference samples, expected within 3 SE, empirical p-values within the binomial CI"""
    from gat_b200 import engine as Engine
    _check_distributions("distribution", tag, lambda meta: Engine.SamplerAnnotator())


@pytest.mark.parametrize("tag", ["plain", "iso"])
def test_shift_sampler_distributions_equivalent_to_reference(ctx, tag):
    """the same equivalence for --sampler=shift: 2000 samples of the reference's SamplerShift(radius=3)
    against 4000 of the GPU's through gat_b200.run"""
    from gat_b200 import engine as Engine
    _check_distributions("distribution_shift", tag,
                         lambda meta: Engine.SamplerShift(radius=meta["shift"][0], extension=meta["shift"][1]))


def _check_distributions(fixture, tag, make_sampler):
    import gat_b200
    from gat_b200 import engine as Engine
    z, meta = G.load_npz(fixture)
    m = meta[tag]
    segments, annotations, workspace, _, _, _, _ = _track_problem(z, tag, m)
    counters = m["counters"]
    n = 4000
    Engine.seed(2026)
    res = gat_b200.run(segments, annotations, workspace, make_sampler(meta), counters_of(counters),
                       Engine.UnconditionalWorkspace(), num_samples=n)
    assert len(res) == len(counters) * len(m["annotation_order"])
    nref = m["num_samples"]
    crit = 1.95 * np.sqrt((n + nref) / (n * nref))             # alpha ~ 0.001 per test
    for r in res:
        ai = m["annotation_order"].index(r.annotation)
        ref = z["%s/samples/%s" % (tag, r.counter)][:, ai].astype(np.float64)
        assert r.observed == z["%s/observed/%s" % (tag, r.counter)][ai]
        mine = r.samples
        assert ks_statistic(mine, ref) < crit, (r.counter, r.annotation, ks_statistic(mine, ref), crit)
        se = np.sqrt(ref.var() / nref + mine.var() / n)
        assert abs(mine.mean() - ref.mean()) <= 3 * se + 1e-9, (r.counter, r.annotation)
        # fold within 3 SE (delta method on the expected value)
        fold_ref = (r.observed + 1.0) / (ref.mean() + 1.0)
        assert abs(r.fold - fold_ref) <= 3 * se * (r.observed + 1.0) / (ref.mean() + 1.0) ** 2 + 1e-9
        # p-values: both estimate the same tail probability
        pref = float(z["%s/pvalue/%s" % (tag, r.counter)][ai])
        p = 0.5 * (pref + r.pvalue)
        sd = np.sqrt(p * (1 - p) * (1.0 / n + 1.0 / nref))
        assert abs(r.pvalue - pref) <= 3.5 * sd + 1.0 / nref, (r.counter, r.annotation, r.pvalue, pref)


def test_run_api_results_and_output(ctx, tmp_path):
    """gat.run()-style call: result objects, counts dump, q-values and the output table"""
    import gat_b200
    from gat_b200 import engine as Engine, io as IO, synthetic
    segments, annotations, workspaces, iso = synthetic.make(
        n_segments=300, n_annotations=5, n_annotation_intervals=400, isochores=True,
        genome=[("chrA", 2000000), ("chrB", 1000000)], isochore_tile=50000, n_isochores=3)
    workspace = synthetic.prepare(segments, annotations, workspaces, iso)
    counters = counters_of(["nucleotide-overlap", "nucleotide-density", "segment-overlap"])
    Engine.seed(5)
    pattern = str(tmp_path / "counts_%s.tsv")
    res = gat_b200.run(segments, annotations, workspace, Engine.SamplerAnnotator(), counters,
                       Engine.UnconditionalWorkspace(), num_samples=200, output_counts_pattern=pattern)
    assert len(res) == 15
    Engine.seed(5)
    res2 = gat_b200.run(segments, annotations, workspace, Engine.SamplerAnnotator(), counters,
                        Engine.UnconditionalWorkspace(), num_samples=200)
    for a, b in zip(res, res2):                                   # same seed -> same samples
        assert np.array_equal(a.samples, b.samples) and a.pvalue == b.pvalue
    for r in res:
        assert r.nsamples == 200 and len(r.samples) == 200
        assert r.expected == pytest.approx(r.samples.mean(), rel=1e-12)
        assert 0 < r.pvalue <= 1 and r.qvalue == 1.0
        if r.counter == "nucleotide-density":
            nuc = [x for x in res if x.counter == "nucleotide-overlap" and x.annotation == r.annotation][0]
            assert r.observed > 0 and r.observed <= nuc.observed
    # counts table round trip (gat/__init__.py:1072-1119)
    back = gat_b200.fromCounts(pattern % "segment-overlap")
    so = [r for r in res if r.counter == "segment-overlap"]
    assert len(back) == len(so)
    for a, b in zip(back, so):
        assert np.array_equal(a.samples, b.samples) and a.pvalue == b.pvalue and a.expected == b.expected

    class Opt(object):
        qvalue_method = "BH"
        qvalue_lambda = None
        qvalue_pi0_method = "smoother"
        output_order = "fold"
        output_tables_pattern = str(tmp_path / "table_%s.tsv")
        stdout = io.StringIO()
    IO.outputResults(res, Opt, Engine.AnnotatorResultExtended.headers, [], 0, {})
    lines = open(str(tmp_path / "table_nucleotide-overlap.tsv")).read().strip().split("\n")
    assert lines[0].split("\t") == Engine.AnnotatorResultExtended.headers and len(lines) == 6
    folds = [float(l.split("\t")[7]) for l in lines[1:]]
    assert folds == sorted(folds)
    # the size / overlap columns come from one batched GPU call per track (Engine.overlapColumns); the
    # per-result host path of the reference's constructor (intersect per result) must give the same row
    for r in res:
        host = Engine.AnnotatorResultExtended(
            r.track, r.annotation, r.counter, r.observed, r.samples, segments[r.track], annotations[r.annotation],
            workspace, stats=dict((k, getattr(r, k)) for k in ("expected", "stddev", "lower95", "upper95", "fold", "pvalue")))
        host.qvalue, host.format_observed = r.qvalue, r.format_observed
        assert str(host) == str(r)
        assert r.overlap_size > 0 and r.overlap_nsegments > 0
    assert all(0 < r.qvalue <= 1 for r in res)
    q = Engine.getQValues([r.pvalue for r in res], method="BH")
    assert np.allclose([r.qvalue for r in res], q)


def test_unaccelerated_options_raise(ctx):
    import gat_b200
    from gat_b200 import engine as Engine

    class Other(object):
        pass

    class Cond(Engine.UnconditionalWorkspace):
        is_conditional = True
    with pytest.raises(NotImplementedError):
        gat_b200.run(None, None, None, Other(), [], Engine.UnconditionalWorkspace())
    with pytest.raises(NotImplementedError):
        gat_b200.run(None, None, None, Engine.SamplerAnnotator(), [], Cond())


def test_operator_protocols_single_calls(ctx, oracle):
    """the reference's per-call protocols: sampler.sample(segments, workspace) and counter(segments,
    annotations, workspace) work on single lists (gat/Engine.pyx:515-517, 1417-1472)"""
    from gat_b200 import engine as Engine
    from gat_b200.segmentlist import SegmentList
    segs = SegmentList(iter=[(100, 200), (1000, 1300), (5000, 5050)], normalize=True)
    ws = SegmentList(iter=[(0, 3000), (4000, 10000)], normalize=True)
    Engine.seed(3)
    out = Engine.SamplerAnnotator().sample(segs, ws)
    assert out.isNormalized and len(out) >= 1
    t = out.clone()
    t.intersect(ws)
    assert t.sum() == 450                                   # exactly the input's workspace coverage
    annos = SegmentList(iter=[(150, 1100), (5040, 6000)], normalize=True)
    for name, cls in Engine.COUNTER_CLASSES.items():
        got = cls()(segs, annos, ws)
        assert got == oracle.counter(name, segs.asarray(), annos.asarray(), len(ws)), name
    assert Engine.SamplerAnnotator().sample(SegmentList(), ws).isEmpty
    # SamplerShift.sample: the moved segments stay inside the workspace and lose no base except by overlap
    moved = Engine.SamplerShift(radius=4).sample(segs, ws)
    assert moved.isNormalized and 0 < moved.sum() <= 450
    inside = moved.clone()
    inside.intersect(ws)
    assert inside.sum() == moved.sum()
    assert Engine.SamplerShift().sample(SegmentList(), ws).isEmpty
    with pytest.raises(ValueError):                          # segment too large for the histogram
        Engine.SamplerAnnotator(bucket_size=1, nbuckets=100).sample(segs, ws)


@pytest.mark.parametrize("with_isochores", [False, True])
def test_cli_from_bed_files(ctx, oracle, tmp_path, with_isochores):
    """gat-run.py-style invocation: BED files in, result table out (--segments/--annotations/--workspace/
    --isochore-file, --counter, --num-samples, --random-seed); observed column checked against the oracle"""
    from gat_b200 import cli, synthetic, io as IO
    import gat_b200
    genome = [("chrA", 1500000), ("chrB", 700000)]
    segments, annotations, workspaces, iso = synthetic.make(
        n_segments=250, n_annotations=4, n_annotation_intervals=300, isochores=with_isochores, genome=genome,
        isochore_tile=100000, n_isochores=2, seed=99)
    files = {}
    for name, coll in (("segments", segments), ("annotations", annotations), ("workspace", workspaces)):
        files[name] = str(tmp_path / (name + ".bed"))
        synthetic.write_bed(coll, files[name], with_tracks=name == "annotations")
    argv = ["gat-run", "--segments=" + files["segments"], "--annotations=" + files["annotations"],
            "--workspace=" + files["workspace"], "--counter=nucleotide-overlap", "--counter=segment-overlap",
            "--num-samples=300", "--random-seed=7", "--output-tables-pattern=" + str(tmp_path / "out_%s.tsv")]
    if with_isochores:
        files["iso"] = str(tmp_path / "iso.bed")
        synthetic.write_bed(iso, files["iso"], with_tracks=True)
        argv.append("--isochore-file=" + files["iso"])
    assert cli.main(argv + ["-v", "0"]) == 0
    # expected observed values: the same files through the host preparation and the oracle's counters
    options, _ = gat_b200.buildParser().parse_args(argv[1:])
    s2, a2, w2, i2 = IO.buildSegments(options)
    ws = IO.applyIsochores(s2, a2, w2, options, i2)
    for counter in ("nucleotide-overlap", "segment-overlap"):
        rows = [l.rstrip("\n").split("\t") for l in open(str(tmp_path / ("out_%s.tsv" % counter)))]
        assert rows[0][:11] == ["track", "annotation", "observed", "expected", "CI95low", "CI95high", "stddev", "fold",
                                "l2fold", "pvalue", "qvalue"]
        assert len(rows) == 5
        for r in rows[1:]:
            want = sum(oracle.counter(counter, s2["merged"][k].asarray(), a2[r[1]][k].asarray(), len(ws[k]))
                       for k in ws.keys())
            assert int(r[2]) == int(want), (counter, r[1])
            assert 0 < float(r[9]) <= 1 and 0 < float(r[10]) <= 1 and float(r[3]) > 0
    # --sampler=shift (scripts/gat-run.py:129-132): same files, same observed column, its own expectation
    first = dict((r[1], r) for r in rows[1:])
    shift_argv = [a for a in argv if not a.startswith("--output-tables-pattern")]
    shift_argv += ["--sampler=shift", "--shift-expansion=3", "--output-tables-pattern=" + str(tmp_path / "shift_%s.tsv")]
    assert cli.main(shift_argv + ["-v", "0"]) == 0
    rows = [l.rstrip("\n").split("\t") for l in open(str(tmp_path / "shift_segment-overlap.tsv"))]
    assert len(rows) == 5
    for r in rows[1:]:
        assert r[2] == first[r[1]][2] and float(r[3]) > 0 and 0 < float(r[9]) <= 1


def test_compare_cli_matches_reference_tables(ctx, tmp_path, monkeypatch):
    """gat_b200.compare (gatb_compare_stats: derived log-ratio samples + column statistics on the GPU) prints the
    rows the reference's scripts/gat-compare.py printed for the same count tables (tests/golden/compare.json)"""
    from gat_b200 import compare
    data = G.load_json("compare")
    monkeypatch.chdir(tmp_path)
    for name, text in data["files"].items():
        with open(name, "w") as f:
            f.write(text)
    for k, case in enumerate(data["cases"]):
        out = "out%i.tsv" % k
        assert compare.main(["gat-compare"] + case["args"] + ["--stdout=" + out]) == 0
        got = [l.rstrip("\n") for l in open(out)]
        assert got[0] == case["table"][0]
        assert sorted(got[1:]) == sorted(case["table"][1:]), case["args"]
        if "--order=annotation" in case["args"]:
            assert got == case["table"]                       # no ties in this order: same sequence too


def test_compare_stats_large_matches_oracle(ctx, oracle):
    """5 000 samples x 40 columns, all 780 pairs in one call: statistics against the oracle's restatement"""
    rng = np.random.default_rng(31)
    S, A = 5000, 40
    lam = rng.uniform(0.5, 60, A)
    m = rng.poisson(lam, size=(S, A)).astype(np.float64)
    obs = rng.poisson(lam * rng.choice([0.5, 1, 2], A)).astype(np.float64)
    base = ctx.column_stats(m.astype(np.uint32), obs, pseudo_count=1.0)
    pairs = [(i, j) for i in range(A) for j in range(i + 1, A)]
    c1, c2 = [p[0] for p in pairs], [p[1] for p in pairs]
    delta = base["fold"][c2] - base["fold"][c1]
    got = ctx.compare_stats(m, None, c1, c2, obs[c1], obs[c2], delta, pseudo_count=1.0)
    for q in range(0, len(pairs), 37):
        i, j = pairs[q]
        d, st = oracle.compare_pair(obs[i], m[:, i], base["fold"][i], obs[j], m[:, j], base["fold"][j], 1.0)
        assert d == delta[q]
        assert got["pvalue"][q] == st.pvalue                                    # counts of x<obs, x==obs: exact
        assert got["lower95"][q] == pytest.approx(st.lower95, rel=1e-12, abs=1e-12)   # log(): <= 1 ulp apart
        assert got["upper95"][q] == pytest.approx(st.upper95, rel=1e-12, abs=1e-12)
        assert got["expected"][q] == pytest.approx(st.expected, rel=1e-10, abs=1e-12)
        assert got["stddev"][q] == pytest.approx(st.stddev, rel=1e-10)
        assert got["fold"][q] == pytest.approx(st.fold, rel=1e-9)
