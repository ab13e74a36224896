"""Pin the CPU oracle (oracle/gat_oracle.c) against the reference: golden vectors produced by the compiled
reference (tests/golden/make_golden.py) and the known answers of the reference's own tests.
Everything is bit-exact (integer work; float64 statistics restated with numpy's summation order)."""
import numpy as np
import pytest

from tests import golden_util as G

COUNTERS = ["nucleotide-overlap", "nucleotide-density", "segment-overlap", "segment-midoverlap",
            "annotation-overlap", "annotation-midoverlap"]


def arr(x):
    return np.array(x, dtype=np.uint32).reshape(-1, 2)


# ------------------------------------------------------------------------------- interval algebra
def test_segmentlist_ops_match_reference(oracle):
    cases = G.load_json("segmentlist")
    assert len(cases) >= 50
    for c in cases:
        na, nb = oracle.normalize(c["a"]), oracle.normalize(c["b"])
        assert np.array_equal(na, arr(c["normalize_a"]))
        assert np.array_equal(nb, arr(c["normalize_b"]))
        for d in (0, 1, 5):
            assert np.array_equal(oracle.merge(c["a"], d), arr(c["merge_%i" % d]))
        assert np.array_equal(oracle.filter_(na, nb), arr(c["filter"]))
        assert np.array_equal(oracle.intersect(na, nb), arr(c["intersect"]))
        assert oracle.total(na) == c["sum_a"]
        assert oracle.overlap_with_segments(na, nb) == c["overlap"]
        assert oracle.intersection_with_segments(na, nb) == c["isect_base"]
        assert oracle.intersection_with_segments(na, nb, "midpoint") == c["isect_mid"]
        if "insertion_point" in c:
            s, e, ip = c["insertion_point"]
            assert oracle.get_insertion_point(na, s, e) == ip
        if "trim_ends" in c:
            t = c["trim_ends"]
            assert np.array_equal(oracle.trim_ends(na, t["pos"], t["size"], t["forward"]), arr(t["result"]))
        for bucket in (0, 1, 4):
            key = "lengthdist_%i" % bucket
            if key not in c:
                continue
            if "error" in c[key]:
                with pytest.raises(ValueError):
                    oracle.length_distribution(na, bucket, 1000)
                continue
            h, bs = oracle.length_distribution(na, bucket, 1000)
            assert bs == c[key]["bucket_size"]
            assert [[int(k), int(h[k])] for k in np.flatnonzero(h)] == c[key]["nonzero"]


def test_reference_unit_test_known_answers(oracle):
    """known answers of the reference's own unit tests (test/test_SegmentList.py)"""
    # testNormalize1..5 (:27-77): overlapping merge, adjacent stay apart, empties vanish
    assert oracle.normalize([(0, 10), (5, 15)]).tolist() == [[0, 15]]
    assert oracle.normalize([(0, 10), (10, 20)]).tolist() == [[0, 10], [10, 20]]
    assert oracle.normalize([(0, 10), (0, 10), (0, 10)]).tolist() == [[0, 10]]
    assert oracle.normalize([(5, 5), (0, 0)]).tolist() == []
    assert oracle.normalize([(0, 100), (5, 10), (20, 30)]).tolist() == [[0, 100]]
    # testMergeAdjacent / testMergeNeighbours (:206-227)
    assert oracle.merge([(0, 10), (10, 20), (20, 30)], 0).tolist() == [[0, 30]]
    assert oracle.merge([(0, 10), (11, 20), (22, 30)], 0).tolist() == [[0, 10], [11, 20], [22, 30]]
    assert oracle.merge([(0, 10), (11, 20), (22, 30)], 1).tolist() == [[0, 20], [22, 30]]
    assert oracle.merge([(0, 10), (11, 20), (22, 30)], 2).tolist() == [[0, 30]]
    # testIntersect / filter (:369-441)
    a = [(0, 10), (10, 20), (30, 40)]
    assert oracle.intersect(a, [(5, 35)]).tolist() == [[5, 10], [10, 20], [30, 35]]
    assert oracle.intersect(a, [(0, 20)]).tolist() == [[0, 10], [10, 20]]      # pieces not re-merged
    assert oracle.filter_(a, [(5, 8)]).tolist() == [[0, 10]]
    assert oracle.filter_(a, [(20, 30)]).tolist() == []
    assert oracle.filter_(a, [(19, 31)]).tolist() == [[10, 20], [30, 40]]
    assert oracle.intersection_with_segments([(0, 10), (20, 30)], [(5, 25)]) == 2
    assert oracle.overlap_with_segments([(0, 10), (20, 30)], [(5, 25)]) == 10


# ---------------------------------------------------------------------------------------- sampler
def test_sampler_annotator_matches_reference_under_numpy_rng(oracle):
    """the restatement driven by numpy's global RNG returns the reference's segments, unit by unit"""
    data = G.load_json("sampler_units")
    units = data["units"]
    assert len(units) >= 100
    ntrims = 0
    for u in units:
        np.random.seed(u["seed"])
        got, info = oracle.sampler_annotator_numpy(u["segments"], u["workspace"], bucket_size=u["bucket_size"])
        assert np.array_equal(got, arr(u["placed"])), u["seed"]
        ntrims += info.ntrims
    assert ntrims > 20
    t = data["too_large"]
    assert t["error"] == "ValueError"
    with pytest.raises(ValueError):
        oracle.sampler_annotator_numpy(t["segments"], t["workspace"], bucket_size=t["bucket_size"], nbuckets=t["nbuckets"])


def test_sampler_invariants(oracle):
    """every sample covers exactly the input's workspace bases and only holds workspace-overlapping
    segments (test/benchmark_gat.py:773-780, 828-837), for the Philox stream too"""
    from tests import helpers
    rng = np.random.default_rng(3)
    for it in range(40):
        segs, ws = helpers.random_unit(rng)
        got, info = oracle.sampler_annotator_philox(segs, ws, seed=5, track=0, unit=it, sample=it)
        want = oracle.total(oracle.intersect(oracle.filter_(segs, ws), ws))
        if info.nunsuccessful < 20:
            assert oracle.total(oracle.intersect(got, ws)) == want
        assert len(oracle.filter_(got, ws)) == len(got)
        assert np.array_equal(oracle.merge(got, 0), got)


def test_philox_known_answer(oracle):
    """Philox4x32-10 known-answer vectors of the Random123 distribution (kat_vectors)"""
    assert oracle.philox4x32_10([0, 0, 0, 0], [0, 0]).tolist() == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert oracle.philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2).tolist() == \
        [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert oracle.philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]).tolist() == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    # bounded integers stay in range and hit both ends
    vals = [oracle.philox_randint(1, 0, 0, 0, t, 0, 3, 10) for t in range(400)]
    assert min(vals) == 3 and max(vals) == 9


# --------------------------------------------------------------------------------------- counters
def test_counters_match_reference(oracle):
    cases = G.load_json("counters")
    for c in cases:
        for name, want in zip(COUNTERS, c["counts"]):
            assert oracle.counter(name, c["segments"], c["annotations"], c["workspace_nsegments"]) == want, name


def test_same_placement_counts_match_reference(oracle):
    """per-sample counts of the reference's OWN placed samples (--output-samples-pattern), all six
    counters, without and with isochores: placed units -> per-contig merge(0) -> counts"""
    z, meta = G.load_npz("run_small")
    for tag in ("plain", "iso"):
        m = meta[tag]
        segments = G.collection(z, tag + "/segments", m["segments"])
        annotations = G.collection(z, tag + "/annotations", m["annotations"])
        workspace = G.dictionary(z, tag + "/workspace", m["workspace"])
        import gat_b200
        atracks, lists, nseg = None, None, None
        for s, placed in enumerate(m["placed"]):
            # contig order = first appearance among the placed unit keys (sample.keys() after fromIsochores)
            contigs, per = [], {}
            for key, segs in placed.items():
                contig = key.split(".")[0]
                if contig not in per:
                    per[contig] = []
                    contigs.append(contig)
                per[contig].extend(segs)
            # units that returned nothing are absent from the dump: take the contig order of the track
            problem = gat_b200.TrackProblem(segments["merged"], workspace)
            contigs = problem.contigs
            atracks, lists, nseg = gat_b200.buildContigAnnotations(annotations, workspace, contigs)
            sample = []
            for c in contigs:
                a = arr(per.get(c, []))
                sample.append(oracle.merge(a, 0) if tag == "iso" else a)
            got = oracle.count_placed(sample, lists, nseg, COUNTERS)
            for ci, counter in enumerate(COUNTERS):
                for ai, anno in enumerate(atracks):
                    want = m["results"][counter][anno]["samples"][s]
                    assert got[ci, ai] == want, (tag, s, counter, anno, got[ci, ai], want)


def _observed_with_oracle(oracle, z, meta, counter="nucleotide-overlap"):
    out = {}
    for track in sorted(set(t for t, _ in meta["segments"])):
        for anno in sorted(set(t for t, _ in meta["annotations"])):
            total = 0
            for key in meta["workspace"]:
                sk = "segments/%s/%s" % (track, key)
                ak = "annotations/%s/%s" % (anno, key)
                if sk not in z.files or ak not in z.files:
                    continue
                total += oracle.counter(counter, G.undelta(z[sk]), G.undelta(z[ak]), 1)
            out["%s|%s" % (track, anno)] = int(total)
    return out


def test_observed_counts_reference_golden_run(oracle):
    """the 28 observed values of the reference's golden output test/data/output_single.tsv"""
    z, meta = G.load_npz("observed_testdata")
    got = _observed_with_oracle(oracle, z, meta)
    assert len(meta["golden"]) == 28
    assert got == meta["golden"]


def test_observed_count_tutorial(oracle):
    """SRF x Jurkat DHS observed overlap 20183 (doc/tutorialIntervalOverlap.rst:103)"""
    z, meta = G.load_npz("observed_tutorial")
    got = _observed_with_oracle(oracle, z, meta)
    assert list(got.values()) == [20183]


# ------------------------------------------------------------------------------------- statistics
def test_statistics_match_reference(oracle):
    cases = G.load_json("stats")
    assert len(cases) > 50
    for c in cases:
        st = oracle.enrichment_statistics(c["observed"], c["samples"], pseudo_count=c["pseudo_count"],
                                          reference_fold=c["ref_fold"])
        assert st.expected == c["expected"]
        assert st.stddev == c["stddev"]
        assert st.fold == c["fold"]
        assert st.pvalue == c["pvalue"], c["row"]
        assert "%6.4f" % st.lower95 == "%6.4f" % c["lower95"]
        assert "%6.4f" % st.upper95 == "%6.4f" % c["upper95"]


def test_pvalue_known_answers_of_reference_tests(oracle):
    """test/test_gat.py:120-129 (p = 0.57) and :272-284 (p = 0.01)"""
    samples = list(range(0, 100))
    st = oracle.enrichment_statistics(56, samples)           # obs above the mean: (100 - 57 + ...)/100
    ref = [c for c in G.load_json("stats") if c.get("origin")]
    assert ref and ref[0]["pvalue"] == oracle.enrichment_statistics(ref[0]["observed"], ref[0]["samples"]).pvalue
    # observed far outside the samples -> 1/nsamples
    assert oracle.enrichment_statistics(1000, samples).pvalue == 0.01
    assert oracle.enrichment_statistics(-5, samples).pvalue == 0.01
    assert 0.0 < st.pvalue <= 1.0


def test_adjust_pvalues_match_reference(oracle):
    for c in G.load_json("qvalues"):
        for method, want in c["adjusted"].items():
            got = oracle.adjust_pvalues(c["pvalues"], method)
            assert np.array_equal(got, np.array(want)), method


def test_sampler_segments_matches_reference_under_numpy_rng(oracle):
    """SamplerSegments.sample (gat/Engine.pyx:653-737): placements in draw order equal the reference's"""
    units = G.load_json("sampler_segments")
    assert len(units) >= 40
    for u in units:
        np.random.seed(u["seed"])
        got = oracle.sampler_segments(u["segments"], u["workspace"], bucket_size=u["bucket_size"])
        assert got.tolist() == u["placed"], u["seed"]
        assert len(got) in (0, len(u["segments"]))


def test_sampler_shift_matches_reference_under_numpy_rng(oracle):
    """SamplerShift.sample (gat/Engine.pyx:998-1111): the normalized sample equals the reference's for the
    default radius, other radii and extensions (fixtures include wrapped shifts, fragmented workspaces and
    segments whose window holds no workspace)"""
    units = G.load_json("sampler_shift")
    assert len(units) >= 60
    nonempty = 0
    for u in units:
        np.random.seed(u["seed"])
        got = oracle.sampler_shift(u["segments"], u["workspace"], radius=u["radius"], extension=u["extension"])
        assert got.tolist() == u["placed"], u["seed"]
        nonempty += len(got) > 0
    assert nonempty >= 50


def test_sampler_shift_invariants_under_philox(oracle):
    """properties of the shift that hold for any RNG, checked on the oracle under the kernel's Philox stream:
    the sample is normalized, lies inside the workspace, never holds more bases than the working segments
    (bases are only lost where two moved segments overlap), is a pure function of (seed, track, unit, sample),
    and a segment whose window holds no workspace drops out (the reference's ignored ValueError)"""
    from tests import helpers
    rng = np.random.default_rng(404)
    for it in range(60):
        segs, ws = helpers.random_unit(rng)
        kw = [dict(radius=2, extension=0), dict(radius=0.5, extension=0), dict(radius=2, extension=600)][it % 3]
        key = (1000 + it, it % 3, 0, 5)
        got = oracle.sampler_shift(segs, ws, philox=key, **kw)
        again = oracle.sampler_shift(segs, ws, philox=key, **kw)
        assert np.array_equal(got, again)
        if len(got):
            assert (got[:, 0] < got[:, 1]).all() and (got[1:, 0] >= got[:-1, 1]).all()       # normalized
            assert oracle.overlap_with_segments(got, ws) == int((got[:, 1] - got[:, 0]).sum())   # inside
        working = oracle.filter_(segs, ws)
        assert int((got[:, 1] - got[:, 0]).sum()) <= int((working[:, 1] - working[:, 0]).sum())
    # extension = 4: a segment is wrapped inside the 4 workspace bases around its midpoint (a fill longer than
    # the local workspace returns the local workspace, gat/SegmentList.pyx:1325-1326); the window of the second
    # segment, [1003, 1007), lies in the workspace gap, so that segment drops out
    ws = np.array([[0, 1000], [5000, 6000]], dtype=np.uint32)
    segs = np.array([[100, 200], [990, 1020]], dtype=np.uint32)
    for sample in range(8):
        got = oracle.sampler_shift(segs, ws, radius=2, extension=4, philox=(1, 0, 0, sample))
        assert got.tolist() == [[148, 152]], got.tolist()


def _parse_counts(text):
    rows = []
    for line in text.splitlines()[1:]:
        track, annotation, observed, counts = line.split("\t")
        rows.append((track, annotation, float(observed), np.array([float(x) for x in counts.split(",")])))
    return rows


def test_compare_matches_reference_script(oracle):
    """gat-compare restatement (oracle.compare_pair on top of the statistics oracle) against the tables the
    reference's scripts/gat-compare.py printed (tests/golden/compare.json): every printed column of every row"""
    data = G.load_json("compare")
    tables = dict((name, _parse_counts(text)) for name, text in data["files"].items())
    stats = {}
    for name, rows in tables.items():
        for track, annotation, observed, samples in rows:
            stats[(name, track, annotation)] = (observed, samples, oracle.enrichment_statistics(observed, samples).fold)
    for case in data["cases"]:
        files = [a for a in case["args"] if not a.startswith("-")]
        pc = 1.0
        for a in case["args"]:
            if a.startswith("--pseudo-count="):
                pc = float(a.split("=")[1])
        want = dict(((r.split("\t")[0], r.split("\t")[1], r.split("\t")[2]), r.split("\t")) for r in case["table"][1:])
        got = []
        if len(files) == 1:
            rows = tables[files[0]]
            for i in range(len(rows)):
                for j in range(i + 1, len(rows)):
                    a, b = stats[(files[0],) + rows[i][:2]], stats[(files[0],) + rows[j][:2]]
                    got.append((rows[i][1], rows[j][1]) + oracle.compare_pair(a[0], a[1], a[2], b[0], b[1], b[2], pc))
        else:
            for x in range(len(files)):
                for y in range(x + 1, len(files)):
                    keys_a = set(r[:2] for r in tables[files[x]])
                    for track, annotation in sorted(keys_a.intersection(r[:2] for r in tables[files[y]])):
                        a, b = stats[(files[x], track, annotation)], stats[(files[y], track, annotation)]
                        got.append((track, annotation) + oracle.compare_pair(a[0], a[1], a[2], b[0], b[1], b[2], pc))
        assert len(got) == len(case["table"]) - 1
        for track, annotation, delta, st in got:
            row = want[(track, annotation, "%6.4f" % delta)]
            assert row[3] == "%6.4f" % st.expected and row[4] == "%6.4f" % st.lower95, (case["args"], row)
            assert row[5] == "%6.4f" % st.upper95 and row[6] == "%6.4f" % st.stddev, (case["args"], row)
            assert row[7] == "%6.4f" % st.fold and row[9] == "%6.4e" % st.pvalue, (case["args"], row)
