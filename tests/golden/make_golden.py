#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ from the REFERENCE ITSELF.

Run in the build container only (needs oracle/_ref, built from /root/reference by
oracle/build_ref.py, and /root/reference for the reference's own test/tutorial data):

    python tests/golden/make_golden.py

Every fixture records inputs and the outputs the compiled reference (AndreasHeger/gat 1.3.6) produced
for them, so that tests can pin the oracle and the CUDA path without the reference being present
(the GPU box has neither /root/reference nor a need for oracle/_ref in the tests).

  segmentlist.json   SegmentList ops on random lists (normalize, merge, filter, intersect, overlap, ...)
  sampler_units.json SamplerAnnotator.sample under numpy.random.seed(seed): units + placed segments
  counters.json      the six Counter classes on random list pairs
  stats.json         AnnotatorResult statistics (expected, stddev, CI95, fold, p) incl. reference folds
  qvalues.json       Stats.adjustPValues (all methods) and Stats.computeQValues
  run_small.npz      gat.run on a small synthetic problem (with and without isochores): placed samples
                     (--output-samples-pattern), per-sample counts, observed counts, result rows
  distribution.npz   sampled-count distributions (2000 samples) for the KS / 3-SE equivalence tests
  observed_testdata.npz  the reference's own golden run test/data/output_single.tsv: prepared
                     intervals + the 28 observed values of that file (checked here against a live run)
  observed_tutorial.npz  tutorial SRF x Jurkat DHS: prepared intervals + published observed 20183
  compare.json       scripts/gat-compare.py on small count tables (one file / several files / options): the
                     input tables and the result tables the reference printed
"""
import gzip
import io
import json
import os
import re
import sys
import tempfile
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REFERENCE = "/root/reference"

from oracle import ref_bench  # noqa: E402

warnings.filterwarnings("ignore")
gat = ref_bench.load()
import gat.Engine as RE  # noqa: E402
import gat.SegmentList as RS  # noqa: E402
import gat.Stats as RStats  # noqa: E402
import gat.IO as RIO  # noqa: E402

COUNTERS = ["nucleotide-overlap", "nucleotide-density", "segment-overlap", "segment-midoverlap",
            "annotation-overlap", "annotation-midoverlap"]


def L(sl):
    return [[int(a), int(b)] for a, b in sl]


def rlist(rng, span, n, maxlen):
    out = []
    for _ in range(n):
        s = int(rng.integers(0, span))
        out.append((s, s + int(rng.integers(1, maxlen + 1))))
    return out


def make_segmentlist(rng):
    cases = []
    for it in range(60):
        span = int(rng.choice([60, 1000, 100000]))
        a = rlist(rng, span, int(rng.integers(0, 25)), int(rng.choice([3, 30, 300])))
        b = rlist(rng, span, int(rng.integers(0, 25)), int(rng.choice([3, 30, 300])))
        if it % 7 == 0 and a:
            a.append((a[0][0], a[0][0]))                      # an empty segment
        c = dict(a=a, b=b)
        na = RS.SegmentList(iter=a, normalize=True)
        nb = RS.SegmentList(iter=b, normalize=True)
        c["normalize_a"] = L(na)
        c["normalize_b"] = L(nb)
        for d in (0, 1, 5):
            m = RS.SegmentList(iter=a)
            m.merge(d)
            c["merge_%i" % d] = L(m)
        f = RS.SegmentList(clone=na)
        f.filter(nb)
        c["filter"] = L(f)
        i = RS.SegmentList(clone=na)
        i.intersect(nb)
        c["intersect"] = L(i)
        c["sum_a"] = int(na.sum())
        c["overlap"] = int(na.overlapWithSegments(nb))
        c["isect_base"] = int(na.intersectionWithSegments(nb))
        c["isect_mid"] = int(na.intersectionWithSegments(nb, mode="midpoint"))
        if len(na):
            pos = int(rng.integers(0, span))
            c["insertion_point"] = [pos, pos + 1, int(na.getInsertionPoint(pos, pos + 1))]
            total = na.sum()
            if total > 1:
                size = int(rng.integers(1, total))
                pos = int(na[int(rng.integers(0, len(na)))][0])
                fw = int(rng.integers(0, 2))
                t = RS.SegmentList(clone=na)
                t.trim_ends(pos, size, fw)
                c["trim_ends"] = dict(pos=pos, size=size, forward=fw, result=L(t))
            for bucket in (0, 1, 4):
                try:
                    h, bs = na.getLengthDistribution(bucket, 1000)
                    nz = np.flatnonzero(h)
                    c["lengthdist_%i" % bucket] = dict(bucket_size=int(bs), nonzero=[[int(k), int(h[k])] for k in nz])
                except ValueError:
                    c["lengthdist_%i" % bucket] = dict(error="ValueError")
        cases.append(c)
    return cases


def make_sampler_units(rng):
    units = []
    for it in range(120):
        nws = int(rng.integers(1, 8))
        span = int(rng.choice([2000, 20000, 1000000, 30000000]))
        pts = np.sort(rng.choice(span, size=2 * nws, replace=False))
        ws = [(int(pts[2 * i]), int(pts[2 * i + 1])) for i in range(nws)]
        segs = rlist(rng, span, int(rng.integers(1, 70)), int(rng.choice([5, 50, 500, 4000])))
        bucket = int(rng.choice([1, 1, 1, 3, 7]))
        sl = RS.SegmentList(iter=segs, normalize=True)
        wl = RS.SegmentList(iter=ws, normalize=True)
        np.random.seed(1000 + it)
        sampler = RE.SamplerAnnotator(bucket_size=bucket, nbuckets=100000)
        out = sampler.sample(sl, wl)
        units.append(dict(segments=L(sl), workspace=L(wl), bucket_size=bucket, seed=1000 + it, placed=L(out)))
    # the reference's error path: a segment too large for nbuckets * bucket_size
    sl = RS.SegmentList(iter=[(0, 5000)], normalize=True)
    wl = RS.SegmentList(iter=[(0, 100000)], normalize=True)
    try:
        RE.SamplerAnnotator(bucket_size=1, nbuckets=1000).sample(sl, wl)
        err = None
    except ValueError as e:
        err = "ValueError"
    return dict(units=units, too_large=dict(segments=L(sl), workspace=L(wl), bucket_size=1, nbuckets=1000, error=err))


def make_sampler_segments(rng):
    """SamplerSegments.sample (gat/Engine.pyx:653-737) under numpy.random.seed: placements in draw order"""
    units = []
    for it in range(40):
        nws = int(rng.integers(1, 6))
        span = int(rng.choice([5000, 200000, 30000000]))
        pts = np.sort(rng.choice(span, size=2 * nws, replace=False))
        ws = [(int(pts[2 * i]), int(pts[2 * i + 1])) for i in range(nws)]
        segs = rlist(rng, span, int(rng.integers(1, 50)), int(rng.choice([5, 80, 900])))
        bucket = int(rng.choice([1, 1, 4]))
        sl = RS.SegmentList(iter=segs, normalize=True)
        wl = RS.SegmentList(iter=ws, normalize=True)
        np.random.seed(5000 + it)
        out = RE.SamplerSegments(bucket_size=bucket, nbuckets=100000).sample(sl, wl)
        units.append(dict(segments=L(sl), workspace=L(wl), bucket_size=bucket, seed=5000 + it, placed=L(out)))
    return units


def make_sampler_shift(rng):
    """SamplerShift.sample (gat/Engine.pyx:998-1111) under numpy.random.seed: default radius, other radii and
    --shift-extension; fragmented workspaces, segments at the contig start (negative shifted starts), windows
    that hold no workspace (the reference drops such a segment: getRandomPosition's ValueError is ignored)"""
    import contextlib
    units = []
    for it in range(60):
        nws = int(rng.integers(1, 9))
        span = int(rng.choice([3000, 200000, 30000000]))
        pts = np.sort(rng.choice(span, size=2 * nws, replace=False))
        ws = [(int(pts[2 * i]), int(pts[2 * i + 1])) for i in range(nws)]
        if it % 5 == 0:
            ws[0] = (0, ws[0][1])
        segs = rlist(rng, span, int(rng.integers(1, 50)), int(rng.choice([5, 80, 900, 4000])))
        kw = [{}, {"radius": float(rng.choice([0.5, 1, 3, 7.5]))}, {"extension": int(rng.choice([10, 101, 1000, 5000]))}][it % 3]
        sl = RS.SegmentList(iter=segs, normalize=True)
        wl = RS.SegmentList(iter=ws, normalize=True)
        np.random.seed(7000 + it)
        with open(os.devnull, "w") as devnull, contextlib.redirect_stderr(devnull):
            out = RE.SamplerShift(**kw).sample(sl, wl)
        units.append(dict(segments=L(sl), workspace=L(wl), radius=kw.get("radius", 2), extension=kw.get("extension", 0),
                          seed=7000 + it, placed=L(out)))
    return units


def counter_objs():
    return [RE.CounterNucleotideOverlap(), RE.CounterNucleotideDensity(), RE.CounterSegmentOverlap(),
            RE.CounterSegmentMidpointOverlap(), RE.CounterAnnotationOverlap(), RE.CounterAnnotationMidpointOverlap()]


def make_counters(rng):
    cases = []
    objs = counter_objs()
    for it in range(80):
        span = int(rng.choice([200, 5000, 1000000]))
        a = RS.SegmentList(iter=rlist(rng, span, int(rng.integers(0, 40)), int(rng.choice([3, 40, 600]))), normalize=True)
        b = RS.SegmentList(iter=rlist(rng, span, int(rng.integers(0, 40)), int(rng.choice([3, 40, 600]))), normalize=True)
        w = RS.SegmentList(iter=rlist(rng, span, int(rng.integers(0, 4)), 50), normalize=True)
        vals = [float(c(a, b, w)) for c in objs]
        cases.append(dict(segments=L(a), annotations=L(b), workspace_nsegments=len(w), counts=vals))
    return cases


def make_stats(rng):
    cases = []

    def one(observed, samples, pseudo=1.0, ref_fold=None):
        class Ref(object):
            pass
        ref = None
        if ref_fold is not None:
            ref = Ref()
            ref.fold = ref_fold
        r = RE.AnnotatorResult("t", "a", "c", observed, samples, reference=ref, pseudo_count=pseudo)
        return dict(observed=float(observed), samples=[float(x) for x in samples], pseudo_count=pseudo,
                    ref_fold=ref_fold, expected=r.expected, stddev=r.stddev, fold=r.fold, pvalue=r.pvalue,
                    lower95=float(str(r).split("\t")[4]), upper95=float(str(r).split("\t")[5]), row=str(r))

    # known answers of the reference's own tests (test/test_gat.py:120-129, 272-284)
    cases.append(dict(one(8, list(range(0, 100))), expect_pvalue=None, origin="test_gat.py testPValue-like"))
    for l in (1, 2, 10, 19, 20, 21, 100, 1000):
        for _ in range(4):
            lam = float(rng.choice([0.5, 5, 50, 5000]))
            samples = rng.poisson(lam, l).tolist()
            obs = int(rng.poisson(lam * float(rng.choice([0.5, 1, 2]))))
            cases.append(one(obs, samples, pseudo=float(rng.choice([1.0, 0.0, 5.0]))))
    cases.append(one(3, [3] * 50))
    cases.append(one(0, [0] * 50))
    cases.append(one(7, [0] * 50))
    for _ in range(10):                         # float samples (nucleotide-density) + truncating comparator
        samples = np.round(rng.random(200) * 5, 3).tolist()
        cases.append(one(float(np.round(rng.random() * 5, 3)), samples))
    for _ in range(8):                          # --null reference fold
        samples = rng.poisson(30, 300).tolist()
        cases.append(one(int(rng.poisson(40)), samples, ref_fold=float(rng.choice([0.5, 1.5, 2.0]))))
    return cases


def make_qvalues(rng):
    cases = []
    for n in (1, 2, 5, 28, 200):
        for kind in range(3):
            if kind == 0:
                p = rng.random(n)
            elif kind == 1:
                p = np.round(rng.random(n), 1)                 # many ties
            else:
                p = np.concatenate([rng.random(n // 2) * 1e-3, rng.random(n - n // 2)])
            p = np.maximum(p, 1e-4)
            c = dict(pvalues=p.tolist(), adjusted={})
            for method in ("BH", "fdr", "bonferroni", "holm", "hochberg", "BY", "none"):
                c["adjusted"][method] = [float(x) for x in RStats.adjustPValues(p.tolist(), method=method)]
            if n >= 28:
                try:
                    r = RStats.computeQValues(p.tolist(), vlambda=np.arange(0, 0.95, 0.05), pi0_method="smoother")
                    c["storey_smoother"] = dict(pi0=float(r.pi0), qvalues=[float(x) for x in r.qvalues])
                except ValueError as e:
                    c["storey_smoother"] = dict(error=str(e))
                r = RStats.computeQValues(p.tolist(), vlambda=0.5)
                c["storey_lambda05"] = dict(pi0=float(r.pi0), qvalues=[float(x) for x in r.qvalues])
            cases.append(c)
    return cases


# ------------------------------------------------------------------------------------------ whole runs
def small_problem(isochores, seed=5):
    """a small synthetic genome through our generator, prepared with OUR host code (tested separately
    against the reference's IO) and converted to reference objects"""
    from gat_b200 import synthetic
    genome = [("chrA", 3000000), ("chrB", 1700000), ("chrC", 900000), ("chrD", 400000)]
    segments, annotations, workspaces, iso = synthetic.make(
        n_segments=400, n_annotations=6, n_annotation_intervals=600, isochores=isochores, genome=genome,
        seed=seed, isochore_tile=50000, n_isochores=3)
    workspace = synthetic.prepare(segments, annotations, workspaces, iso)
    return segments, annotations, workspace


def dump_collection(prefix, coll, store):
    names = []
    for track, vv in coll.items():
        for key, s in vv.items():
            names.append([track, key])
            store["%s/%s/%s" % (prefix, track, key)] = np.asarray(s.asarray() if hasattr(s, "asarray") else L(s), dtype=np.uint32).reshape(-1, 2)
    return names


def make_run_small():
    store, meta = {}, {}
    for tag, iso in (("plain", False), ("iso", True)):
        segments, annotations, workspace = small_problem(iso)
        meta[tag] = dict(segments=dump_collection(tag + "/segments", segments, store),
                         annotations=dump_collection(tag + "/annotations", annotations, store))
        meta[tag]["workspace"] = []
        for key, s in workspace.items():
            meta[tag]["workspace"].append(key)
            store["%s/workspace/%s" % (tag, key)] = s.asarray()
        nsamples = 6
        with tempfile.TemporaryDirectory() as tmp:
            spat = os.path.join(tmp, "samples_%s.bed")
            cpat = os.path.join(tmp, "counts_%s.tsv")
            results = ref_bench.run_full(segments, annotations, workspace, COUNTERS, nsamples, seed=3,
                                         output_samples_pattern=spat, output_counts_pattern=cpat)
            # placed segments: per sample, per unit key (gat/__init__.py:549-559)
            placed = []
            cur = None
            for line in open(os.path.join(tmp, "samples_merged.bed")):
                if line.startswith("track"):
                    cur = {}
                    placed.append(cur)
                    continue
                k, s, e = line.rstrip("\n").split("\t")
                cur.setdefault(k, []).append([int(s), int(e)])
            meta[tag]["placed"] = placed
            counts = {}
            for c in COUNTERS:
                rows = [l.rstrip("\n").split("\t") for l in open(os.path.join(tmp, "counts_%s.tsv" % c))][1:]
                # NB: the reference prints nucleotide-density with %i (gat/__init__.py:1082-1085): use the
                # result objects for exact values
                counts[c] = {}
        res = {}
        for r in results:
            res.setdefault(r.counter, {})[r.annotation] = dict(observed=float(r.observed),
                                                               samples=[float(x) for x in r.samples],
                                                               expected=r.expected, stddev=r.stddev, fold=r.fold,
                                                               pvalue=r.pvalue, row=str(r))
        meta[tag]["results"] = res
        meta[tag]["num_samples"] = nsamples
    store["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "run_small.npz"), **store)


def make_distribution(shift=None, name="distribution"):
    """reference distributions for the statistical-equivalence tests (KS, 3 SE, binomial CI); with `shift` =
    (radius, extension) the reference samples with SamplerShift"""
    store, meta = {}, {}
    for tag, iso, counters in (("plain", False, ["nucleotide-overlap", "segment-overlap"]),
                               ("iso", True, ["segment-overlap", "nucleotide-overlap"])):
        segments, annotations, workspace = small_problem(iso, seed=11)
        meta[tag] = dict(segments=dump_collection(tag + "/segments", segments, store),
                         annotations=dump_collection(tag + "/annotations", annotations, store), workspace=[])
        for key, s in workspace.items():
            meta[tag]["workspace"].append(key)
            store["%s/workspace/%s" % (tag, key)] = s.asarray()
        nsamples = 2000
        results = ref_bench.run_full(segments, annotations, workspace, counters, nsamples, seed=17,
                                     **({"shift": shift} if shift else {}))
        meta[tag]["counters"] = counters
        meta[tag]["num_samples"] = nsamples
        annos = sorted(set(r.annotation for r in results))
        meta[tag]["annotation_order"] = annos
        for c in counters:
            m = np.zeros((nsamples, len(annos)))
            obs = np.zeros(len(annos))
            pv = np.zeros(len(annos))
            for r in results:
                if r.counter == c:
                    m[:, annos.index(r.annotation)] = r.samples
                    obs[annos.index(r.annotation)] = r.observed
                    pv[annos.index(r.annotation)] = r.pvalue
            store["%s/samples/%s" % (tag, c)] = m.astype(np.uint32)
            store["%s/observed/%s" % (tag, c)] = obs
            store["%s/pvalue/%s" % (tag, c)] = pv
    store["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    if shift:
        meta["shift"] = list(shift)
        store["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **store)


# ------------------------------------------------------------------------------- reference's own data
class Opt(object):
    pass


def ref_options(**kw):
    parser = gat.buildParser()
    options, _ = parser.parse_args([])
    options.output_stats = []
    options.output_bed = []
    for k, v in kw.items():
        setattr(options, k, v)
    return options


def prepared_from_files(segment_files, annotation_files, workspace_files, ignore_segment_tracks):
    options = ref_options(segment_files=segment_files, annotation_files=annotation_files,
                          workspace_files=workspace_files, ignore_segment_tracks=ignore_segment_tracks)
    segments, annotations, workspaces, isochores = RIO.buildSegments(options)
    workspace = RIO.applyIsochores(segments, annotations, workspaces, options, isochores)
    return segments, annotations, workspace


def delta(arr):
    """(n,2) sorted intervals -> (n,2) uint32 [start - previous start, length]: compresses ~3x better;
    tests/golden_util.py:undelta inverts it"""
    a = np.asarray(arr, dtype=np.int64).reshape(-1, 2)
    out = np.empty_like(a)
    out[:, 0] = np.diff(a[:, 0], prepend=0)
    out[:, 1] = a[:, 1] - a[:, 0]
    return out.astype(np.uint32)


def pack(prefix, coll, store):
    """reference collection -> delta-coded uint32 arrays; returns [[track, key], ...]"""
    names = []
    for track, vv in coll.items():
        for key, s in vv.items():
            names.append([track, key])
            store["%s/%s/%s" % (prefix, track, key)] = delta(L(s))
    return names


def make_observed_testdata():
    d = os.path.join(REFERENCE, "test", "data")
    segments, annotations, workspace = prepared_from_files(
        [os.path.join(d, "segments_single.bed.gz")], [os.path.join(d, "annotations.bed.gz")],
        [os.path.join(d, "workspace.bed.gz")], ignore_segment_tracks=False)
    golden = {}
    for line in open(os.path.join(d, "output_single.tsv")):
        if line.startswith("#") or line.startswith("track"):
            continue
        f = line.split("\t")
        golden["%s|%s" % (f[0], f[1])] = int(f[2])
    # live check: the compiled reference reproduces its 2013 golden file exactly
    live = RE.computeCounts(RE.CounterNucleotideOverlap(), sum, segments, annotations, workspace,
                            RE.UnconditionalWorkspace())
    for track, r in live.items():
        for anno, v in r.items():
            assert golden["%s|%s" % (track, anno)] == int(v), (track, anno, v)
    assert len(golden) == 28
    store = {}
    meta = dict(golden=golden, segments=pack("segments", segments, store),
                annotations=pack("annotations", annotations, store), workspace=[],
                source="test/data/output_single.tsv (observed column), ignore_segment_tracks=False")
    for key, s in workspace.items():
        meta["workspace"].append(key)
        store["workspace/%s" % key] = delta(L(s))
    store["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "observed_testdata.npz"), **store)


def make_observed_tutorial():
    d = os.path.join(REFERENCE, "tutorial", "TutorialIntervalOverlap")
    segments, annotations, workspace = prepared_from_files(
        [os.path.join(d, "srf.hg19.bed.gz")], [os.path.join(d, "jurkat.hg19.dhs.bed.gz")],
        [os.path.join(d, "contigs.bed.gz")], ignore_segment_tracks=True)
    live = RE.computeCounts(RE.CounterNucleotideOverlap(), sum, segments, annotations, workspace,
                            RE.UnconditionalWorkspace())
    (track, r), = live.items()
    (anno, v), = r.items()
    assert int(v) == 20183, v            # doc/tutorialIntervalOverlap.rst:103
    store = {}
    meta = dict(golden={"%s|%s" % (track, anno): 20183}, segments=pack("segments", segments, store),
                annotations=pack("annotations", annotations, store), workspace=[],
                source="doc/tutorialIntervalOverlap.rst:103 (SRF x Jurkat DHS, contig workspace)",
                published=dict(expected=246.5650, stddev=105.5933, fold=81.5301, pvalue=1.0e-3,
                               ci95=[96, 444], nsamples=1000))
    for key, s in workspace.items():
        meta["workspace"].append(key)
        store["workspace/%s" % key] = delta(L(s))
    store["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "observed_tutorial.npz"), **store)


def make_prep_isochores(rng):
    """the reference's own IO (buildSegments + applyIsochores, gat/IO.py:88-293) on small BED files WITH an
    isochore file, default and --truncate-segments-to-workspace: every prepared list, for the host preparation
    of gat_b200 (f1) to reproduce.  The BED texts travel in the fixture."""
    from gat_b200 import synthetic
    genome = [("chrA", 1500000), ("chrB", 700000), ("chrC", 300000)]
    segments, annotations, workspaces, iso = synthetic.make(
        n_segments=400, n_annotations=5, n_annotation_intervals=500, isochores=True, genome=genome,
        isochore_tile=50000, n_isochores=3, seed=int(rng.integers(1, 1 << 30)))
    out = {"files": {}, "cases": []}
    with tempfile.TemporaryDirectory() as d:
        argv = []
        for name, coll, flag in (("segments", segments, "--segments"), ("annotations", annotations, "--annotations"),
                                 ("workspace", workspaces, "--workspace"), ("iso", iso, "--isochore-file")):
            path = os.path.join(d, name + ".bed")
            synthetic.write_bed(coll, path, with_tracks=name in ("annotations", "iso"))
            out["files"][name + ".bed"] = open(path).read()
            argv.append("%s=%s" % (flag, path))
        for extra in ([], ["--truncate-segments-to-workspace"]):
            ropt, _ = gat.buildParser().parse_args(argv + extra)
            rs, ra, rw, ri = RIO.buildSegments(ropt)
            rws = RIO.applyIsochores(rs, ra, rw, ropt, ri,
                                     truncate_segments_to_workspace=ropt.truncate_segments_to_workspace)
            case = {"extra": extra, "workspace": dict((k, L(rws[k])) for k in rws.keys())}
            for name, coll in (("segments", rs), ("annotations", ra)):
                case[name] = dict((t, dict((k, L(coll[t][k])) for k in coll[t].keys())) for t in coll.tracks)
            out["cases"].append(case)
    return out


def make_output_tables(rng):
    """the reference's IO.outputResults (gat/IO.py:457-538) on the results of one of its own small runs: q-values
    over all results, one table per counter, every --output-order, BH and storey q-values, with and without
    annotation descriptions.  Stored: what is needed to rebuild the result objects, and the tables' text."""
    segments, annotations, workspace = small_problem(False, seed=int(rng.integers(1, 1 << 30)))
    counters = ["nucleotide-overlap", "segment-overlap"]
    results = ref_bench.run_full(segments, annotations, workspace, counters, 200, seed=5)
    out = {"results": [], "cases": [], "headers": list(RE.AnnotatorResultExtended.headers),
           "workspace_size": int(sum(int(workspace[k].sum()) for k in workspace.keys()))}
    for r in results:
        row = str(r).split("\t")
        out["results"].append(dict(track=r.track, annotation=r.annotation, counter=r.counter, observed=float(r.observed),
                                   expected=r.expected, stddev=r.stddev, fold=r.fold, pvalue=r.pvalue,
                                   lower95=float(row[4]), upper95=float(row[5]), tail=row[11:]))
    descriptions = dict((a, ["desc of %s" % a, "x"]) for a in sorted(set(r.annotation for r in results))[:3])
    with tempfile.TemporaryDirectory() as d:
        for order in ("track", "observed", "annotation", "fold", "pvalue", "qvalue"):
            for method, with_desc in (("BH", False), ("storey", False), ("BH", True)):
                class O(object):
                    pass
                O.qvalue_method, O.qvalue_lambda, O.qvalue_pi0_method = method, None, "smoother"
                O.output_order = order
                O.output_tables_pattern = os.path.join(d, "t_%s.tsv")
                O.stdout = sys.stdout
                rs = list(results)
                RIO.outputResults(rs, O, RE.AnnotatorResultExtended.headers, ["description", "extra"] if with_desc else [],
                                  2 if with_desc else 0, descriptions if with_desc else {})
                tables = dict((c, open(os.path.join(d, "t_%s.tsv" % c)).read()) for c in counters)
                out["cases"].append(dict(order=order, method=method, with_desc=with_desc, tables=tables))
    out["descriptions"] = descriptions
    return out


def make_parser_options(rng):
    """every option of the reference's gat.buildParser() (gat/__init__.py:190-427): flags, action, type, default"""
    p = gat.buildParser()
    out = {}
    for o in p._get_all_options():
        if o.dest:
            d = p.defaults.get(o.dest)
            out[o.dest] = dict(flags=sorted(o._long_opts + o._short_opts), action=o.action, type=o.type,
                               default=d if isinstance(d, (int, float, str, bool, list, type(None))) else str(d),
                               choices=list(o.choices) if o.choices else None)
    return out


def make_compare(rng):
    """scripts/gat-compare.py run by the reference on small count tables: within one file and between files"""
    import runpy

    def counts_text(tracks, annos, S, scale, zero_every=0):
        lines = ["track\tannotation\tobserved\tcounts\n"]
        for t in tracks:
            for i, a in enumerate(annos):
                lam = scale * (i + 1) * 2.5
                samples = rng.poisson(lam, S)
                if zero_every:
                    samples[::zero_every] = 0
                obs = int(rng.poisson(lam * rng.choice([0.4, 1.0, 2.5])))
                lines.append("%s\t%s\t%i\t%s\n" % (t, a, obs, ",".join(str(x) for x in samples)))
        return "".join(lines)

    files = {"a.tsv": counts_text(["merged"], ["x1", "x2", "x3", "x4", "x5"], 200, 1.0, zero_every=17),
             "b.tsv": counts_text(["merged"], ["x1", "x2", "x3", "x5", "x9"], 200, 1.6),
             "c.tsv": counts_text(["merged", "other"], ["x1", "x2", "x5"], 200, 0.7, zero_every=5)}
    script = os.path.join(ROOT, "oracle", "_ref", "scripts", "gat-compare.py")
    cases = []
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as d:
        os.chdir(d)
        try:
            for name, text in files.items():
                with open(name, "w") as f:
                    f.write(text)
            for args in (["a.tsv"], ["a.tsv", "b.tsv"], ["a.tsv", "b.tsv", "c.tsv"],
                         ["a.tsv", "--pseudo-count=0.5", "--order=annotation", "--qvalue-method=bonferroni"]):
                out, old_out, old_argv = io.StringIO(), sys.stdout, sys.argv
                sys.stdout, sys.argv = out, ["gat-compare.py"] + args + ["--log=/dev/null"]
                try:
                    runpy.run_path(script, run_name="__main__")
                except SystemExit:
                    pass
                finally:
                    sys.stdout, sys.argv = old_out, old_argv
                table = [l for l in out.getvalue().splitlines() if l and not l.startswith("#")]
                cases.append({"args": args, "table": table})
        finally:
            os.chdir(cwd)
    return {"files": files, "cases": cases}


def main():
    # fixtures added after the first generation have their own seeds and can be (re)made alone:
    #   python tests/golden/make_golden.py sampler_segments
    extra = {"sampler_segments": (make_sampler_segments, 20260102), "compare": (make_compare, 20260103),
             "sampler_shift": (make_sampler_shift, 20260104), "prep_isochores": (make_prep_isochores, 20260105),
             "output_tables": (make_output_tables, 20260106), "parser_options": (make_parser_options, 20260107)}
    if "distribution_shift" in sys.argv[1:]:
        make_distribution(shift=(3.0, 0), name="distribution_shift")
        print("wrote distribution_shift")
        return
    only = [a for a in sys.argv[1:] if a in extra]
    for name in (only or list(extra)):
        fn, seed = extra[name]
        with open(os.path.join(HERE, name + ".json"), "w") as f:
            json.dump(fn(np.random.default_rng(seed)), f)
        print("wrote", name)
    if only:
        return
    rng = np.random.default_rng(20260101)
    for name, fn in (("segmentlist", make_segmentlist), ("sampler_units", make_sampler_units),
                     ("counters", make_counters), ("stats", make_stats), ("qvalues", make_qvalues)):
        data = fn(rng)
        with open(os.path.join(HERE, name + ".json"), "w") as f:
            json.dump(data, f)
        print("wrote", name)
    make_run_small()
    print("wrote run_small")
    make_distribution()
    print("wrote distribution")
    make_observed_testdata()
    print("wrote observed_testdata")
    make_observed_tutorial()
    print("wrote observed_tutorial")


if __name__ == "__main__":
    main()
