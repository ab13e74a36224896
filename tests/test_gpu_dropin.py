"""The drop-in claim, exercised: the UNMODIFIED reference -- compiled under oracle/_ref by
oracle/build_ref.py -- runs its own gat.run() with the binding of INTEGRATION.md section 2 (integration/_b200.py,
ctypes against libgat_b200.so, nothing of the gat_b200 package) installed as UnconditionalSampler.sample and
Engine.computeCounts.  Inputs: the reference's golden run (test/data, prepared intervals in
tests/golden/observed_testdata.npz) and the tutorial (tests/golden/observed_tutorial.npz), plus the small runs whose
sampled distributions the reference produced itself (tests/golden/distribution.npz).

  observed            bit-exact against the reference's own numbers (28 golden values, 20183)
  expected / fold     within 3 standard errors of the reference's own samples
  same matrix         the reference-driven run and gat_b200.run() see identical sample columns for one seed
"""
import importlib.util
import os

import numpy as np
import pytest

from tests import golden_util as G

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ref():
    """(reference package `gat`, the binding module), binding installed"""
    from oracle import ref_bench
    if not ref_bench.available():
        pytest.fail("oracle/_ref is not built (python oracle/build_ref.py): the drop-in test needs the compiled reference")
    gat = ref_bench.load()
    spec = importlib.util.spec_from_file_location("gat_b200_binding", os.path.join(ROOT, "integration", "_b200.py"))
    binding = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(binding)
    uninstall = binding.install(gat)
    yield gat, binding
    uninstall()


def _to_ref(gat, coll=None, dictionary=None):
    from oracle import ref_bench
    if coll is not None:
        return ref_bench.ref_collection(gat, coll, coll.name or "x")
    return ref_bench.ref_dictionary(gat, dictionary)


def test_reference_run_on_the_library_golden_testdata(ctx, ref):
    """the reference's own golden run (test/data/output_single.tsv): 28 observed values bit-exact through
    gatb_count_lists, a full gat.run() with 200 samples on top of gatb_run"""
    gat, binding = ref
    z, meta = G.load_npz("observed_testdata")
    segments = G.collection(z, "segments", meta["segments"], coded=True)
    annotations = G.collection(z, "annotations", meta["annotations"], coded=True)
    workspace = G.dictionary(z, "workspace", meta["workspace"], coded=True)
    rs, ra = _to_ref(gat, segments), _to_ref(gat, annotations)
    rw = _to_ref(gat, dictionary=workspace)
    binding.seed(11)
    results = gat.run(rs, ra, rw, gat.Engine.SamplerAnnotator(bucket_size=1, nbuckets=100000),
                      [gat.Engine.CounterNucleotideOverlap()], gat.Engine.UnconditionalWorkspace(), num_samples=200)
    got = dict(("%s|%s" % (r.track, r.annotation), r) for r in results)
    assert len(meta["golden"]) == 28 and sorted(got) == sorted(meta["golden"])
    for key, observed in meta["golden"].items():
        r = got[key]
        assert int(r.observed) == int(observed), key
        assert r.nsamples == 200 and 0 < r.pvalue <= 1 and r.expected >= 0
        assert type(r).__name__ == "AnnotatorResultExtended"        # the reference's own result objects


def test_reference_run_on_the_library_tutorial(ctx, ref):
    """doc/tutorialIntervalOverlap.rst: observed 20183 bit-exact; expected / stddev / fold agree with the published
    1000-sample table (expected 246.565, stddev 105.59, fold 81.53) within Monte-Carlo error"""
    gat, binding = ref
    z, meta = G.load_npz("observed_tutorial")
    segments = G.collection(z, "segments", meta["segments"], coded=True)
    annotations = G.collection(z, "annotations", meta["annotations"], coded=True)
    workspace = G.dictionary(z, "workspace", meta["workspace"], coded=True)
    binding.seed(3)
    results = gat.run(_to_ref(gat, segments), _to_ref(gat, annotations), _to_ref(gat, dictionary=workspace),
                      gat.Engine.SamplerAnnotator(bucket_size=1, nbuckets=100000),
                      [gat.Engine.CounterNucleotideOverlap()], gat.Engine.UnconditionalWorkspace(), num_samples=500)
    assert len(results) == 1
    r, pub = results[0], meta["published"]
    assert int(r.observed) == 20183
    se = np.hypot(pub["stddev"] / np.sqrt(1000), pub["stddev"] / np.sqrt(500))
    assert abs(r.expected - pub["expected"]) < 4 * se, r.expected
    assert abs(r.stddev - pub["stddev"]) / pub["stddev"] < 0.2
    assert abs(r.fold - pub["fold"]) / pub["fold"] < 0.1
    assert r.pvalue == pytest.approx(1.0 / 500)
    # the row prints through the reference's own formatting code
    assert str(r).split("\t")[:3] == [r.track, r.annotation, "20183"]


@pytest.mark.parametrize("tag", ["plain", "iso"])
def test_reference_run_equals_package_run_and_reference_distribution(ctx, ref, tag):
    """one seed, two drivers: the reference's gat.run() over the binding and gat_b200.run() produce the SAME sample
    columns (same library calls underneath), and the expectation agrees with the 2000 samples the reference drew
    with its own sampler (tests/golden/distribution.npz) within 3 SE"""
    import gat_b200
    from gat_b200 import engine as Engine
    gat, binding = ref
    z, meta = G.load_npz("distribution")
    m = meta[tag]
    segments = G.collection(z, tag + "/segments", m["segments"])
    annotations = G.collection(z, tag + "/annotations", m["annotations"])
    workspace = G.dictionary(z, tag + "/workspace", m["workspace"])
    names = ["nucleotide-overlap", "segment-overlap"]
    S = 400
    binding.seed(21)
    ref_results = gat.run(_to_ref(gat, segments), _to_ref(gat, annotations), _to_ref(gat, dictionary=workspace),
                          gat.Engine.SamplerAnnotator(bucket_size=1, nbuckets=100000),
                          [gat.Engine.CounterNucleotideOverlap(), gat.Engine.CounterSegmentOverlap()],
                          gat.Engine.UnconditionalWorkspace(), num_samples=S)
    Engine.seed(21)
    own = gat_b200.run(segments, annotations, workspace, Engine.SamplerAnnotator(bucket_size=1, nbuckets=100000),
                       [Engine.COUNTER_CLASSES[n]() for n in names], Engine.UnconditionalWorkspace(), num_samples=S)
    own = dict(((r.track, r.annotation, r.counter), r) for r in own)
    assert len(ref_results) == len(own) > 0
    for r in ref_results:
        o = own[(r.track, r.annotation, r.counter)]
        assert r.observed == o.observed
        assert np.array_equal(np.asarray(r.samples), o.samples), (r.annotation, r.counter)
        assert r.pvalue == o.pvalue and r.expected == pytest.approx(o.expected, rel=1e-12)
    # against the reference's own sampler: tests/golden/distribution.npz holds its per-sample counts
    ref_samples = z[tag + "/samples/nucleotide-overlap"]        # [n_samples][n_annotations]
    for r in ref_results:
        if r.counter != "nucleotide-overlap":
            continue
        col = ref_samples[:, m["annotation_order"].index(r.annotation)].astype(np.float64)
        se = np.hypot(col.std() / np.sqrt(len(col)), np.asarray(r.samples).std() / np.sqrt(S))
        assert abs(col.mean() - r.expected) <= 3.5 * se + 1e-9, (r.annotation, col.mean(), r.expected)
