"""Host-side SegmentList: the container the reference passes around (gat/SegmentList.pyx), backed by a
numpy (n,2) uint32 array so it can be flattened to the CSR layout of the C ABI without copies.

Only input preparation runs here (load -> normalize -> filter/intersect with the workspace -> split by
isochore; gat/IO.py:88-293).  That is one-off O(input) work outside the simulation loop; the per-sample
interval algebra of the hot path (sort/merge/intersect/overlap inside SamplerAnnotator.sample and the
Counter classes) runs in the CUDA kernels.  Method names and semantics follow the reference so its
tests read the same (tests/test_oracle_golden.py and tests/test_host_logic.py pin it to the reference's results).
"""
import numpy as np

_EMPTY = np.zeros((0, 2), dtype=np.uint32)


class SegmentList(object):
    """list of half-open segments [start, end) on one contig (gat/SegmentList.pyx:145-331)."""

    __slots__ = ("_a", "_normalized")

    def __init__(self, iter=None, normalize=False, clone=None, array=None, allocate=0):
        if clone is not None:
            self._a = clone._a.copy()
            self._normalized = clone._normalized
        elif array is not None:
            self._a = np.ascontiguousarray(array, dtype=np.uint32).reshape(-1, 2)
            self._normalized = len(self._a) == 0
        elif iter is not None:
            data = list(iter)
            for s, e in data:
                if s < 0 or e < 0:
                    raise OverflowError("can't convert negative value to Position")
            self._a = np.array(data, dtype=np.uint32).reshape(-1, 2)
            self._normalized = len(self._a) == 0
        else:
            self._a = _EMPTY.copy()
            self._normalized = True
        if normalize:
            self.normalize()

    # ------------------------------------------------------------------ container protocol
    def __len__(self):
        return len(self._a)

    def __iter__(self):
        for s, e in self._a:
            yield (int(s), int(e))

    def __getitem__(self, i):
        s, e = self._a[i]
        return (int(s), int(e))

    def asList(self):
        return [(int(s), int(e)) for s, e in self._a]

    def asarray(self):
        """the (n,2) uint32 array (no copy)"""
        return self._a

    def __str__(self):
        return str(self.asList())

    def __eq__(self, other):
        return isinstance(other, SegmentList) and np.array_equal(self._a, other._a)

    @property
    def isNormalized(self):
        return self._normalized

    @property
    def isEmpty(self):
        return len(self._a) == 0

    def clone(self):
        return SegmentList(clone=self)

    def clear(self):
        self._a = _EMPTY.copy()
        self._normalized = True

    # ------------------------------------------------------------------ building
    def add(self, start, end):
        assert start <= end, "attempting to add invalid segment %i-%i" % (start, end)
        self._a = np.concatenate([self._a, np.array([[start, end]], dtype=np.uint32)])
        self._normalized = False

    def extend(self, other):
        """append the segments of *other*; list is not normalized afterwards (:488-513)"""
        self._a = np.concatenate([self._a, other._a])
        self._normalized = False
        return self

    def sort(self):
        """sort by start (:478-486; ties in arbitrary order)"""
        if len(self._a):
            self._a = self._a[np.argsort(self._a[:, 0].astype(np.int32), kind="stable")]

    # ------------------------------------------------------------------ normalization
    def _merge(self, distance, adjacent):
        a = self._a
        if len(a) == 0:
            self._normalized = True
            return
        a = a[a[:, 0] != a[:, 1]]                       # drop empty
        if len(a) == 0:
            self._a = _EMPTY.copy()
            self._normalized = True
            return
        a = a[np.argsort(a[:, 0].astype(np.int32), kind="stable")]
        start = a[:, 0].astype(np.int64)
        end = a[:, 1].astype(np.int64)
        cmax = np.maximum.accumulate(end)
        prev = np.empty_like(cmax)
        prev[0] = -1 << 40
        prev[1:] = cmax[:-1]
        if adjacent:
            head = (start - distance) > prev            # merge(): joins when start - distance <= max_end
        else:
            head = start >= prev                        # normalize(): adjacent segments stay apart
        head[0] = True
        idx = np.flatnonzero(head)
        last = np.append(idx[1:] - 1, len(a) - 1)
        out = np.empty((len(idx), 2), dtype=np.uint32)
        out[:, 0] = start[idx]
        out[:, 1] = cmax[last]
        self._a = out
        self._normalized = True

    def normalize(self):
        """merge overlapping segments, remove empty ones; adjacent segments are kept (:697-754)"""
        self._merge(0, adjacent=False)

    def merge(self, distance=0):
        """merge overlapping segments and those at most *distance* apart; distance 0 joins adjacent
        segments (:756-816)"""
        self._merge(int(distance), adjacent=True)

    def check(self):
        """raise ValueError unless sorted, non-empty segments without overlap (:818-851)"""
        a = self._a
        if len(a) == 0:
            self._normalized = True
            return True
        if np.any(a[:, 0] >= a[:, 1]):
            raise ValueError("empty/invalid segment in segmentlist")
        if np.any(a[:-1, 0] > a[1:, 0]):
            raise ValueError("segment list is not sorted")
        if np.any(a[:-1, 1] > a[1:, 0]):
            raise ValueError("segment overlap")
        self._normalized = True
        return True

    # ------------------------------------------------------------------ queries
    def sum(self):
        """total length (:1607-1616)"""
        a = self._a
        return int((a[:, 1].astype(np.int64) - a[:, 0].astype(np.int64)).sum())

    def counts(self):
        return len(self._a)

    def max(self):
        return int(self._a[:, 1].max()) if len(self._a) else 0

    def min(self):
        return int(self._a[:, 0].min()) if len(self._a) else 0

    def largest(self):
        if len(self._a) == 0:
            raise ValueError("largest segment from empty list")
        lens = self._a[:, 1].astype(np.int64) - self._a[:, 0].astype(np.int64)
        i = int(np.argmax(lens))
        return self[i]

    def _overlap_range(self, other):
        """for every segment of self: [j1, j2) = the segments of (normalized) other overlapping it"""
        b = other._a
        j1 = np.searchsorted(b[:, 1], self._a[:, 0], side="right")     # first other with end > start
        j2 = np.searchsorted(b[:, 0], self._a[:, 1], side="left")      # first other with start >= end
        return j1, np.maximum(j1, j2)

    def filter(self, other):
        """keep the segments overlapping *other* by >= 1 base, untruncated (:1401-1467)"""
        if other is self:
            self.clear()
            return
        if len(self._a) == 0:
            return
        if len(other._a) == 0:
            self._a = _EMPTY.copy()
            return
        if self._normalized and len(other._a) == 1 and self._a[0, 0] >= other._a[0, 0] \
                and self._a[-1, 1] <= other._a[0, 1]:
            return          # every segment lies inside the single piece of other
        j1, j2 = self._overlap_range(other)
        self._a = self._a[j2 > j1]

    def intersect(self, other):
        """truncate to the parts shared with *other*; pieces are not re-merged (:1469-1549)"""
        assert self._normalized, "intersection of a non-normalized list"
        assert other._normalized, "intersection with non-normalized list"
        if other is self or len(self._a) == 0:
            return
        if len(other._a) == 0:
            self._a = _EMPTY.copy()
            return
        if len(other._a) == 1 and self._a[0, 0] >= other._a[0, 0] and self._a[-1, 1] <= other._a[0, 1]:
            return          # one piece that holds the whole (sorted) list, e.g. a whole-contig workspace: unchanged
        j1, j2 = self._overlap_range(other)
        cnt = j2 - j1
        total = int(cnt.sum())
        if total == 0:
            self._a = _EMPTY.copy()
            return
        rep = np.repeat(np.arange(len(self._a)), cnt)
        first = np.repeat(np.cumsum(cnt) - cnt, cnt)
        oj = np.repeat(j1, cnt) + (np.arange(total) - first)
        out = np.empty((total, 2), dtype=np.uint32)
        out[:, 0] = np.maximum(self._a[rep, 0], other._a[oj, 0])
        out[:, 1] = np.minimum(self._a[rep, 1], other._a[oj, 1])
        self._a = out

    def overlapWithSegments(self, other):
        """number of bases shared with *other* (:1026-1076)"""
        assert self._normalized, "intersection from non-normalized list"
        assert other._normalized, "intersection with non-normalized list"
        if other is self:
            return self.sum()
        t = self.clone()
        t.intersect(other)
        return t.sum()

    def intersectionWithSegments(self, other, mode="base"):
        """number of segments overlapping *other*; mode 'midpoint' tests the midpoint against the first
        overlapping segment of other only (:1078-1146)"""
        assert self._normalized, "intersection from non-normalized list"
        assert other._normalized, "intersection with non-normalized list"
        if other is self:
            return self.sum()
        if len(self._a) == 0 or len(other._a) == 0:
            return 0
        j1, j2 = self._overlap_range(other)
        hit = j2 > j1
        if mode != "midpoint":
            return int(hit.sum())
        a = self._a[hit]
        o = other._a[j1[hit]]
        mid = a[:, 0].astype(np.int64) + (a[:, 1].astype(np.int64) - a[:, 0].astype(np.int64)) // 2
        return int(((o[:, 0] <= mid) & (mid < o[:, 1])).sum())

    def getLengthDistribution(self, bucket_size=0, nbuckets=100000):
        """histogram of ceil(length / bucket_size) (:1148-1184)"""
        assert bucket_size >= 0, "bucket_size is 0"
        assert nbuckets > 0, "nbuckets is 0"
        import math
        lens = self._a[:, 1].astype(np.int64) - self._a[:, 0].astype(np.int64)
        if bucket_size == 0:
            largest = self.largest()
            bucket_size = int(math.ceil((largest[1] - largest[0]) / float(nbuckets)))
        idx = ((lens + bucket_size - 1) / float(bucket_size)).astype(np.int64)
        if len(idx) and idx.max() >= nbuckets:
            raise ValueError("segment too large: increase nbuckets (%i) or bucket_size (%i)" %
                             (nbuckets, bucket_size))
        return np.bincount(idx, minlength=nbuckets).astype(np.int64), bucket_size

    def truncate(self, start, end):
        """restrict to [start, end)"""
        a = self._a
        keep = (a[:, 1] > start) & (a[:, 0] < end)
        a = a[keep].copy()
        a[:, 0] = np.maximum(a[:, 0], start)
        a[:, 1] = np.minimum(a[:, 1], end)
        self._a = a
