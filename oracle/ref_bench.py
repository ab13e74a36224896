"""Time the reference's own CPU implementation of the hot path.  TEST / BENCH INFRASTRUCTURE ONLY.

Feeds the compiled reference (oracle/_ref, built by oracle/build_ref.py) the same prepared intervals as
the GPU path and times exactly the function gat_b200 replaces:
UnconditionalSampler.sample() (gat/__init__.py:704-778), i.e. the per-sample loop of computeSample
(placement of every unit + counting against every annotation), with `num_threads` = 0 (one core, the
only statistically valid mode of the reference) or > 0 (its multiprocessing.Pool path).

Used by bench.py (`cpu_baseline` and `--impl reference`) and by tests/golden/make_golden.py.
"""
import os
import sys
import time
import warnings

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(_HERE, "_ref")


def available():
    return os.path.isdir(os.path.join(REF, "gat"))


def load():
    """import the reference package from oracle/_ref (never from gat_b200/)"""
    if not available():
        raise ImportError("oracle/_ref is not built (python oracle/build_ref.py)")
    if REF not in sys.path:
        sys.path.insert(0, REF)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import gat
        import gat.Engine
        import gat.SegmentList
    return gat


def ref_segmentlist(gat, arr):
    """(n,2) uint32 array -> reference SegmentList via its unpickling constructor (a memcpy,
    gat/SegmentList.pyx:262-290) -- normalized flag set"""
    a = np.ascontiguousarray(arr, dtype=np.uint32).reshape(-1, 2)
    if len(a) == 0:
        return gat.SegmentList.SegmentList()
    return gat.SegmentList.SegmentList(unreduce=(len(a), len(a), 1, 10000, None, a.tobytes()))


def ref_collection(gat, coll, name):
    """gat_b200 IntervalCollection -> reference IntervalCollection"""
    out = gat.Engine.IntervalCollection(name=name)
    for track, vv in coll.items():
        for key, s in vv.items():
            out.add(track, key, ref_segmentlist(gat, s.asarray()))
    return out


def ref_dictionary(gat, d):
    out = gat.Engine.IntervalDictionary()
    for key, s in d.items():
        out.add(key, ref_segmentlist(gat, s.asarray()))
    return out


def ref_counters(gat, names):
    m = {"nucleotide-overlap": "CounterNucleotideOverlap", "nucleotide-density": "CounterNucleotideDensity",
         "segment-overlap": "CounterSegmentOverlap", "segment-midoverlap": "CounterSegmentMidpointOverlap",
         "annotation-overlap": "CounterAnnotationOverlap", "annotation-midoverlap": "CounterAnnotationMidpointOverlap"}
    return [getattr(gat.Engine, m[n])() for n in names]


def time_sampling(segments, annotations, workspace, counters, num_samples, num_threads=0, seed=1,
                  bucket_size=1, nbuckets=100000, track="merged"):
    """-> (seconds, counts_per_track) for UnconditionalSampler.sample on one track"""
    gat = load()
    rs = ref_collection(gat, segments, "segments")
    ra = ref_collection(gat, annotations, "annotations")
    rw = ref_dictionary(gat, workspace)
    sampler = gat.Engine.SamplerAnnotator(bucket_size=bucket_size, nbuckets=nbuckets)
    cs = ref_counters(gat, counters)
    np.random.seed(seed)
    outer = gat.UnconditionalSampler(num_samples, gat.Engine.Samples(), None, sampler,
                                     gat.Engine.UnconditionalWorkspace(), cs, {}, num_threads=num_threads)
    t0 = time.perf_counter()
    counts = outer.sample(track, None, cs, rs[track], ra, rw, {})
    dt = time.perf_counter() - t0
    return dt, counts


def run_full(segments, annotations, workspace, counters, num_samples, seed=1, **kwargs):
    """gat.run() of the reference on prepared collections -> list of AnnotatorResultExtended"""
    gat = load()
    rs = ref_collection(gat, segments, "segments")
    ra = ref_collection(gat, annotations, "annotations")
    rw = ref_dictionary(gat, workspace)
    np.random.seed(seed)
    shift = kwargs.pop("shift", None)           # (radius, extension): the reference's SamplerShift instead
    if shift is not None:
        sampler = gat.Engine.SamplerShift(radius=shift[0], extension=shift[1])
    else:
        sampler = gat.Engine.SamplerAnnotator(bucket_size=kwargs.pop("bucket_size", 1),
                                              nbuckets=kwargs.pop("nbuckets", 100000))
    return gat.run(rs, ra, rw, sampler,
                   ref_counters(gat, counters), gat.Engine.UnconditionalWorkspace(),
                   num_samples=num_samples, **kwargs)
