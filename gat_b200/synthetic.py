"""Deterministic hg19-shaped synthetic inputs for tests and bench.py.

24 contigs with the hg19 lengths, one workspace segment per contig, log-normal segment and annotation
lengths with genome-uniform starts, optional 100 kb GC-isochore tiles with 8 labels.  The generator is
`numpy.random.default_rng(seed)` with data seed 20260101, so every implementation is fed the same
bytes; `write_bed` dumps them for the reference's gat-run.py.
"""
import numpy as np

from . import engine as Engine
from .segmentlist import SegmentList

HG19 = [("chr1", 249250621), ("chr2", 243199373), ("chr3", 198022430), ("chr4", 191154276),
        ("chr5", 180915260), ("chr6", 171115067), ("chr7", 159138663), ("chr8", 146364022),
        ("chr9", 141213431), ("chr10", 135534747), ("chr11", 135006516), ("chr12", 133851895),
        ("chr13", 115169878), ("chr14", 107349540), ("chr15", 102531392), ("chr16", 90354753),
        ("chr17", 81195210), ("chr18", 78077248), ("chr19", 59128983), ("chr20", 63025520),
        ("chr21", 48129895), ("chr22", 51304566), ("chrX", 155270560), ("chrY", 59373566)]

DATA_SEED = 20260101


def _intervals(rng, n, mu, sigma, lo, hi, genome):
    """n intervals: length clip(lognormal), genome-uniform start, clipped at the contig end;
    returns {contig: (m,2) int64 array}"""
    names = [g[0] for g in genome]
    sizes = np.array([g[1] for g in genome], dtype=np.int64)
    cum = np.concatenate([[0], np.cumsum(sizes)])
    length = np.clip(np.rint(rng.lognormal(mu, sigma, n)), lo, hi).astype(np.int64)
    pos = rng.integers(0, cum[-1], n)
    c = np.searchsorted(cum, pos, side="right") - 1
    start = pos - cum[c]
    end = np.minimum(start + length, sizes[c])
    out = {}
    for ci, name in enumerate(names):
        m = c == ci
        if m.any():
            out[name] = np.stack([start[m], end[m]], axis=1)
    return out


def make(n_segments=10000, n_annotations=50, n_annotation_intervals=20000, isochores=False,
         genome=HG19, seed=DATA_SEED, isochore_tile=100000, n_isochores=8):
    """-> (segments, annotations, workspaces, isochores) as un-prepared IntervalCollections, i.e. what
    IO.buildSegments returns after normalisation (isochores is None unless requested)."""
    rng = np.random.default_rng(seed)
    segments = Engine.IntervalCollection("segments")
    for contig, arr in _intervals(rng, n_segments, np.log(300.0), 0.6, 50, 5000, genome).items():
        segments.add("merged", contig, SegmentList(array=arr.astype(np.uint32)))
    segments.normalize()
    annotations = Engine.IntervalCollection("annotations")
    for a in range(n_annotations):
        for contig, arr in _intervals(rng, n_annotation_intervals, np.log(1000.0), 0.8, 100, 50000, genome).items():
            annotations.add("anno%03i" % a, contig, SegmentList(array=arr.astype(np.uint32)))
    annotations.normalize()
    workspaces = Engine.IntervalCollection("workspaces")
    for contig, size in genome:
        workspaces.add("collapsed", contig, SegmentList(array=np.array([[0, size]], dtype=np.uint32), normalize=True))
    iso = None
    if isochores:
        iso = Engine.IntervalCollection("isochores")
        per = dict(("gc%i" % i, {}) for i in range(n_isochores))
        for contig, size in genome:
            starts = np.arange(0, size, isochore_tile, dtype=np.int64)
            ends = np.minimum(starts + isochore_tile, size)
            labels = rng.integers(0, n_isochores, len(starts))
            for i in range(n_isochores):
                m = labels == i
                if m.any():
                    per["gc%i" % i][contig] = np.stack([starts[m], ends[m]], axis=1)
        for name, contigs in per.items():
            for contig, arr in contigs.items():
                iso.add(name, contig, SegmentList(array=arr.astype(np.uint32)))
        iso.normalize()
        iso.intersect(workspaces["collapsed"])
    return segments, annotations, workspaces, iso


class _Options(object):
    truncate_segments_to_workspace = False


def prepare(segments, annotations, workspaces, isochores=None):
    """IO.applyIsochores with default options -> workspace IntervalDictionary (inputs modified in place)"""
    from . import io as IO
    return IO.applyIsochores(segments, annotations, workspaces, _Options(), isochores)


def write_bed(collection, filename, with_tracks=True):
    """dump a collection as BED (track name in column 4) for the reference's gat-run.py"""
    with open(filename, "w") as f:
        for track, vv in collection.items():
            for contig, s in vv.items():
                a = s.asarray()
                for i in range(len(a)):
                    if with_tracks:
                        f.write("%s\t%i\t%i\t%s\n" % (contig, a[i, 0], a[i, 1], track))
                    else:
                        f.write("%s\t%i\t%i\n" % (contig, a[i, 0], a[i, 1]))
