"""ctypes front-end of the CPU oracle (oracle/gat_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; gat_b200/ never does (tests/test_no_oracle_in_product.py enforces it).

Segment lists are numpy arrays of shape (n, 2), dtype uint32, rows = [start, end).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

COUNTERS = ["nucleotide-overlap", "nucleotide-density", "segment-overlap",
            "segment-midoverlap", "annotation-overlap", "annotation-midoverlap"]
COUNTER_ID = {name: i for i, name in enumerate(COUNTERS)}

(SLOT_LEN, SLOT_JITTER, SLOT_WS_R, SLOT_WS_P, SLOT_TRIM_R, SLOT_TRIM_P, SLOT_TRIM_DIR) = range(7)

RANDINT_FN = ctypes.CFUNCTYPE(ctypes.c_int64, ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_int64)
TURN_FN = ctypes.CFUNCTYPE(None, ctypes.c_void_p)


class SampleInfo(ctypes.Structure):
    _fields_ = [("ltotal", ctypes.c_int32), ("true_remaining", ctypes.c_int32),
                ("nunsuccessful", ctypes.c_int32), ("nturns", ctypes.c_uint32),
                ("nplaced", ctypes.c_uint32), ("ncheckpoints", ctypes.c_uint32),
                ("ntrims", ctypes.c_uint32), ("bucket_size", ctypes.c_uint32)]


class Stats(ctypes.Structure):
    _fields_ = [("observed", ctypes.c_double), ("expected", ctypes.c_double),
                ("stddev", ctypes.c_double), ("lower95", ctypes.c_double),
                ("upper95", ctypes.c_double), ("fold", ctypes.c_double),
                ("pvalue", ctypes.c_double), ("qvalue", ctypes.c_double),
                ("nsamples", ctypes.c_uint32)]


class PhiloxCtx(ctypes.Structure):
    _fields_ = [("seed", ctypes.c_uint64), ("track", ctypes.c_uint32), ("unit", ctypes.c_uint32),
                ("sample", ctypes.c_uint32), ("turn", ctypes.c_uint32)]


def build(force=False):
    """compile liboracle.so with gcc (seconds)."""
    src = os.path.join(_HERE, "gat_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        vp, sz, u32, i32, u64, dbl = (ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint32, ctypes.c_int32,
                                       ctypes.c_uint64, ctypes.c_double)
        L.go_searchsorted_u32.restype = ctypes.c_long
        L.go_searchsorted_u32.argtypes = [vp, sz, u32]
        for name in ("go_normalize",):
            getattr(L, name).restype = sz
            getattr(L, name).argtypes = [vp, sz]
        L.go_sort.restype = None
        L.go_sort.argtypes = [vp, sz]
        L.go_merge.restype = sz
        L.go_merge.argtypes = [vp, sz, i32]
        for name in ("go_filter", "go_intersect"):
            getattr(L, name).restype = sz
            getattr(L, name).argtypes = [vp, sz, vp, sz, vp]
        L.go_sum.restype = u32
        L.go_sum.argtypes = [vp, sz]
        L.go_overlap_with_segments.restype = u32
        L.go_overlap_with_segments.argtypes = [vp, sz, vp, sz]
        L.go_intersection_with_segments.restype = u32
        L.go_intersection_with_segments.argtypes = [vp, sz, vp, sz, ctypes.c_int]
        L.go_get_insertion_point.restype = ctypes.c_int
        L.go_get_insertion_point.argtypes = [vp, sz, u64]      # go_seg by value == 8 bytes in a register
        L.go_trim_ends.restype = ctypes.c_int
        L.go_trim_ends.argtypes = [vp, sz, u32, u32, ctypes.c_int]
        L.go_length_distribution.restype = u32
        L.go_length_distribution.argtypes = [vp, sz, u32, u32, vp]
        L.go_philox4x32_10.restype = None
        L.go_philox4x32_10.argtypes = [vp, vp, vp]
        L.go_philox_begin.restype = None
        L.go_philox_begin.argtypes = [vp, u64, u32, u32, u32]
        L.go_philox_randint.restype = ctypes.c_int64
        L.go_philox_randint.argtypes = [vp, ctypes.c_int, ctypes.c_int64, ctypes.c_int64]
        L.go_philox_next_turn.restype = None
        L.go_philox_next_turn.argtypes = [vp]
        L.go_sampler_annotator.restype = ctypes.c_long
        L.go_sampler_annotator.argtypes = [vp, sz, vp, sz, u32, u32, vp, vp, vp, vp, sz, vp]
        L.go_counter.restype = dbl
        L.go_counter.argtypes = [ctypes.c_int, vp, sz, vp, sz, sz]
        L.go_count_placed.restype = None
        L.go_count_placed.argtypes = [ctypes.c_int, ctypes.c_int, vp, vp, vp, vp, vp, ctypes.c_int, vp, vp]
        L.go_compute_sample_philox.restype = ctypes.c_int
        L.go_compute_sample_philox.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, ctypes.c_int,
                                               vp, vp, vp, vp, vp, vp, vp, u32, u32, u64, u32, u32,
                                               ctypes.c_int, vp, vp, vp, vp, sz]
        L.go_enrichment_statistics.restype = ctypes.c_int
        L.go_enrichment_statistics.argtypes = [dbl, vp, sz, ctypes.c_int, dbl, dbl, vp]
        L.go_two_sided_pvalue.restype = dbl
        L.go_two_sided_pvalue.argtypes = [vp, sz, dbl, dbl]
        _lib = L
    return _lib


def as_segs(x):
    """-> C-contiguous (n,2) uint32 array (copy)."""
    a = np.array(x, dtype=np.uint32).reshape(-1, 2)
    return np.ascontiguousarray(a)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


# ---------------------------------------------------------------------------------- interval algebra
def normalize(x):
    a = as_segs(x)
    return a[:lib().go_normalize(_p(a), len(a))].copy()


def merge(x, distance=0):
    a = as_segs(x)
    return a[:lib().go_merge(_p(a), len(a), distance)].copy()


def filter_(x, other):
    a, b = as_segs(x), as_segs(other)
    out = np.zeros((max(len(a), 1), 2), dtype=np.uint32)
    return out[:lib().go_filter(_p(a), len(a), _p(b), len(b), _p(out))].copy()


def intersect(x, other):
    a, b = as_segs(x), as_segs(other)
    out = np.zeros((len(a) + len(b) + 1, 2), dtype=np.uint32)
    return out[:lib().go_intersect(_p(a), len(a), _p(b), len(b), _p(out))].copy()


def total(x):
    a = as_segs(x)
    return int(lib().go_sum(_p(a), len(a)))


def overlap_with_segments(x, other):
    a, b = as_segs(x), as_segs(other)
    return int(lib().go_overlap_with_segments(_p(a), len(a), _p(b), len(b)))


def intersection_with_segments(x, other, mode="base"):
    a, b = as_segs(x), as_segs(other)
    return int(lib().go_intersection_with_segments(_p(a), len(a), _p(b), len(b), int(mode == "midpoint")))


def get_insertion_point(x, start, end):
    a = as_segs(x)
    return int(lib().go_get_insertion_point(_p(a), len(a), (int(end) << 32) | int(start)))


def trim_ends(x, pos, size, forward):
    a = as_segs(x)
    rc = lib().go_trim_ends(_p(a), len(a), pos, size, int(forward))
    if rc != 0:
        raise AssertionError("trimming more than the total length")
    return a


def length_distribution(x, bucket_size=0, nbuckets=100000):
    a = as_segs(x)
    hist = np.zeros(nbuckets, dtype=np.int64)
    b = lib().go_length_distribution(_p(a), len(a), bucket_size, nbuckets, _p(hist))
    if b == 0:
        raise ValueError("segment too large: increase nbuckets or bucket_size")
    return hist, int(b)


# ------------------------------------------------------------------------------------------- sampler
def _numpy_randint_callback():
    """draw from numpy's process-global legacy RNG, in call order, as the reference does
    (gat/Engine.pyx:299,326,420,433,620)."""
    def cb(ctx, slot, lo, hi):
        return int(np.random.randint(lo, hi))
    return RANDINT_FN(cb)


def sampler_annotator_numpy(segments, workspace, bucket_size=1, nbuckets=100000, cap=None):
    """SamplerAnnotator.sample driven by numpy.random (seed it first) -- equals the reference."""
    a, w = as_segs(segments), as_segs(workspace)
    cap = cap or (4 * len(a) + 1024)
    out = np.zeros((cap, 2), dtype=np.uint32)
    info = SampleInfo()
    cb = _numpy_randint_callback()
    n = lib().go_sampler_annotator(_p(a), len(a), _p(w), len(w), bucket_size, nbuckets,
                                   ctypes.cast(cb, ctypes.c_void_p), None, None,
                                   _p(out), cap, ctypes.addressof(info))
    if n == -2:
        raise ValueError("segment too large: increase nbuckets or bucket_size")
    if n < 0:
        raise RuntimeError("oracle sampler failed: %i" % n)
    return out[:n].copy(), info


def sampler_annotator_philox(segments, workspace, seed, track, unit, sample,
                             bucket_size=1, nbuckets=100000, cap=None):
    """SamplerAnnotator.sample driven by the Philox stream of the CUDA kernel."""
    L = lib()
    a, w = as_segs(segments), as_segs(workspace)
    cap = cap or (4 * len(a) + 1024)
    out = np.zeros((cap, 2), dtype=np.uint32)
    info = SampleInfo()
    ctx = PhiloxCtx()
    L.go_philox_begin(ctypes.addressof(ctx), seed, track, unit, sample)
    n = L.go_sampler_annotator(_p(a), len(a), _p(w), len(w), bucket_size, nbuckets,
                               ctypes.cast(L.go_philox_randint, ctypes.c_void_p),
                               ctypes.cast(L.go_philox_next_turn, ctypes.c_void_p),
                               ctypes.addressof(ctx), _p(out), cap, ctypes.addressof(info))
    if n == -2:
        raise ValueError("segment too large: increase nbuckets or bucket_size")
    if n < 0:
        raise RuntimeError("oracle sampler failed: %i" % n)
    return out[:n].copy(), info


def sampler_segments(segments, workspace, bucket_size=1, nbuckets=100000, philox=None):
    """SamplerSegments.sample (gat/Engine.pyx:653-737): driven by numpy.random (seed it first) when
    philox is None, else by the Philox stream philox=(seed, track, unit, sample)."""
    L = lib()
    L.go_sampler_segments.restype = ctypes.c_long
    L.go_sampler_segments.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t,
                                      ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p,
                                      ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
    a, w = as_segs(segments), as_segs(workspace)
    cap = len(a) + 8
    out = np.zeros((cap, 2), dtype=np.uint32)
    if philox is None:
        cb = _numpy_randint_callback()
        n = L.go_sampler_segments(_p(a), len(a), _p(w), len(w), bucket_size, nbuckets,
                                  ctypes.cast(cb, ctypes.c_void_p), None, None, _p(out), cap)
    else:
        ctx = PhiloxCtx()
        L.go_philox_begin(ctypes.addressof(ctx), *philox)
        n = L.go_sampler_segments(_p(a), len(a), _p(w), len(w), bucket_size, nbuckets,
                                  ctypes.cast(L.go_philox_randint, ctypes.c_void_p),
                                  ctypes.cast(L.go_philox_next_turn, ctypes.c_void_p),
                                  ctypes.addressof(ctx), _p(out), cap)
    if n == -2:
        raise ValueError("segment too large: increase nbuckets or bucket_size")
    if n < 0:
        raise RuntimeError("oracle sampler failed: %i" % n)
    return out[:n].copy()


def sampler_shift(segments, workspace, radius=2, extension=0, philox=None, cap=None):
    """SamplerShift.sample (gat/Engine.pyx:998-1111): driven by numpy.random (seed it first) when philox is
    None, else by the Philox stream philox=(seed, track, unit, sample).  ValueError where the reference
    raises it (no workspace within the shift area of a segment: numpy.random.randint(0, 0))."""
    L = lib()
    L.go_sampler_shift.restype = ctypes.c_long
    L.go_sampler_shift.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t,
                                   ctypes.c_double, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p,
                                   ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
    a, w = as_segs(segments), as_segs(workspace)
    cap = cap or (len(a) * (min(len(w), 64) + 2) + 64)
    out = np.zeros((cap, 2), dtype=np.uint32)
    if philox is None:
        cb = _numpy_randint_callback()
        n = L.go_sampler_shift(_p(a), len(a), _p(w), len(w), float(radius), int(extension),
                               ctypes.cast(cb, ctypes.c_void_p), None, None, _p(out), cap)
    else:
        ctx = PhiloxCtx()
        L.go_philox_begin(ctypes.addressof(ctx), *philox)
        n = L.go_sampler_shift(_p(a), len(a), _p(w), len(w), float(radius), int(extension),
                               ctypes.cast(L.go_philox_randint, ctypes.c_void_p),
                               ctypes.cast(L.go_philox_next_turn, ctypes.c_void_p),
                               ctypes.addressof(ctx), _p(out), cap)
    if n == -5:
        raise ValueError("low >= high")
    if n < 0:
        raise RuntimeError("oracle sampler failed: %i" % n)
    return out[:n].copy()


def set_sampler_kind(kind, radius=2, extension=0):
    """sampler used by compute_sample_philox: 'annotator' (default), 'segments' or 'shift'"""
    L = lib()
    L.go_set_sampler_kind.restype = None
    L.go_set_sampler_kind.argtypes = [ctypes.c_int]
    L.go_set_shift_params.restype = None
    L.go_set_shift_params.argtypes = [ctypes.c_double, ctypes.c_int32]
    L.go_set_sampler_kind({"annotator": 0, "segments": 1, "shift": 2}[kind])
    L.go_set_shift_params(float(radius), int(extension))


def philox4x32_10(ctr, key):
    c = np.array(ctr, dtype=np.uint32)
    k = np.array(key, dtype=np.uint32)
    o = np.zeros(4, dtype=np.uint32)
    lib().go_philox4x32_10(_p(c), _p(k), _p(o))
    return o


def philox_randint(seed, track, unit, sample, turn, slot, lo, hi):
    ctx = PhiloxCtx(seed, track, unit, sample, turn)
    return int(lib().go_philox_randint(ctypes.addressof(ctx), slot, lo, hi))


# ------------------------------------------------------------------------------------------ counters
def counter(name, segments, annotations, workspace_nsegments=0):
    a, b = as_segs(segments), as_segs(annotations)
    return float(lib().go_counter(COUNTER_ID[name], _p(a), len(a), _p(b), len(b), workspace_nsegments))


def _csr(lists):
    off = np.zeros(len(lists) + 1, dtype=np.uint64)
    for i, x in enumerate(lists):
        off[i + 1] = off[i] + len(x)
    if len(lists) and off[-1] > 0:
        data = np.concatenate([as_segs(x) for x in lists], axis=0)
    else:
        data = np.zeros((0, 2), dtype=np.uint32)
    return off, np.ascontiguousarray(data)


def count_placed(placed, annotations, cws_nseg, counters):
    """placed: list over C contigs of segment arrays; annotations: list over A of lists over C."""
    C, A = len(placed), len(annotations)
    poff, pdat = _csr(placed)
    aoff, adat = _csr([annotations[a][c] for a in range(A) for c in range(C)])
    cw = np.array(cws_nseg, dtype=np.uint32)
    cid = np.array([COUNTER_ID[c] for c in counters], dtype=np.int32)
    out = np.zeros((len(cid), A), dtype=np.float64)
    lib().go_count_placed(C, A, _p(poff), _p(pdat), _p(aoff), _p(adat), _p(cw), len(cid), _p(cid), _p(out))
    return out


def compute_sample_philox(unit_contig, unit_segments, unit_workspace, annotations, cws_nseg, counters,
                          seed, track, sample, has_isochores=False, bucket_size=1, nbuckets=100000,
                          return_placed=False, cap=None):
    """one Monte-Carlo sample (gat/__init__.py:494-591) under the Philox stream."""
    U = len(unit_contig)
    A = len(annotations)
    C = len(cws_nseg)
    uc = np.array(unit_contig, dtype=np.int32)
    soff, sdat = _csr(unit_segments)
    woff, wdat = _csr(unit_workspace)
    aoff, adat = _csr([annotations[a][c] for a in range(A) for c in range(C)])
    cw = np.array(cws_nseg, dtype=np.uint32)
    cid = np.array([COUNTER_ID[c] for c in counters], dtype=np.int32)
    out = np.zeros((len(cid), A), dtype=np.float64)
    unit_cap = int(cap or 0)
    cap = int(cap * U if cap else (4 * len(sdat) + 64 * U + 1024))
    poff = np.zeros(C + 1, dtype=np.uint64)
    pdat = np.zeros((cap, 2), dtype=np.uint32)
    lib().go_set_unit_cap.restype = None
    lib().go_set_unit_cap.argtypes = [ctypes.c_size_t]
    lib().go_set_unit_cap(unit_cap)
    rc = lib().go_compute_sample_philox(U, C, A, _p(uc), int(has_isochores), _p(soff), _p(sdat),
                                        _p(woff), _p(wdat), _p(aoff), _p(adat), _p(cw),
                                        bucket_size, nbuckets, seed, track, sample,
                                        len(cid), _p(cid), _p(out), _p(poff), _p(pdat), cap)
    if rc != 0:
        raise RuntimeError("oracle compute_sample failed: %i" % rc)
    if return_placed:
        placed = [pdat[int(poff[c]):int(poff[c + 1])].copy() for c in range(C)]
        return out, placed
    return out


# -------------------------------------------------------------------------------------------- stats
def enrichment_statistics(observed, samples, pseudo_count=1.0, reference_fold=None):
    s = np.ascontiguousarray(np.array(samples, dtype=np.float64))
    st = Stats()
    rc = lib().go_enrichment_statistics(float(observed), _p(s), len(s), int(reference_fold is not None),
                                        float(reference_fold or 0.0), float(pseudo_count),
                                        ctypes.addressof(st))
    if rc != 0:
        raise ValueError("oracle stats failed: %i" % rc)
    return st


def adjust_pvalues(pvalues, method="BH"):
    """gat/Stats.py:192-258 adjustPValues (numpy restatement; R p.adjust port)."""
    p = np.array(pvalues, dtype=np.float64)
    n = lp = len(p)
    if method == "fdr":
        method = "BH"
    if n <= 1:
        return p
    if method == "bonferroni":
        p0 = n * p
    elif method == "holm":
        i = np.arange(lp)
        o = np.argsort(p)
        ro = np.argsort(o)
        p0 = np.maximum.accumulate((n - i) * p[o])[ro]
    elif method == "hochberg":
        i = np.arange(0, lp)[::-1]
        o = np.argsort(1 - p)
        ro = np.argsort(o)
        p0 = np.minimum.accumulate((n - i) * p[o])[ro]
    elif method == "BH":
        i = np.arange(1, lp + 1)[::-1]
        o = np.argsort(1 - p)
        ro = np.argsort(o)
        p0 = np.minimum.accumulate(float(n) / i * p[o])[ro]
    elif method == "BY":
        i = np.arange(1, lp + 1)[::-1]
        o = np.argsort(1 - p)
        ro = np.argsort(o)
        q = np.sum(1.0 / np.arange(1, n + 1))
        p0 = np.minimum.accumulate(q * float(n) / i * p[o])[ro]
    elif method == "none":
        p0 = p
    else:
        raise NotImplementedError(method)
    return np.minimum(p0, np.ones(len(p0)))


def compare_pair(observed1, samples1, fold1, observed2, samples2, fold2, pseudo_count=1.0):
    """scripts/gat-compare.py:218-241 (== :300-323): the relative fold change of two results.
    -> (observed_delta_fold, Stats of AnnotatorResult(observed_delta, sampled_delta, pseudo_count=0))"""
    s1 = np.asarray(samples1, dtype=np.float64)
    s2 = np.asarray(samples2, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        fc1 = observed1 / (s1 + pseudo_count)
        fc2 = observed2 / (s2 + pseudo_count)
        fc1 = fc1 + 0.0001
        fc2 = fc2 + 0.0001
        delta = fold2 - fold1
        sampled = np.log(fc1 / fc2) + delta
    return 0.0 + delta, enrichment_statistics(0.0 + delta, sampled, pseudo_count=0.0)
