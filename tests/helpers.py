"""shared generators for the parity tests (seeded; no reference needed at run time)"""
import numpy as np


def normalize(arr):
    """numpy restatement of SegmentList.normalize for building test inputs"""
    from gat_b200.segmentlist import SegmentList
    s = SegmentList(array=np.asarray(arr, dtype=np.uint32).reshape(-1, 2), normalize=True)
    return s.asarray().copy()


def random_list(rng, span, n, maxlen):
    if n == 0:
        return np.zeros((0, 2), dtype=np.uint32)
    s = rng.integers(0, span, n)
    l = rng.integers(1, maxlen + 1, n)
    return normalize(np.stack([s, s + l], axis=1))


def random_workspace(rng, span, npieces):
    pts = np.sort(rng.choice(span, size=2 * npieces, replace=False))
    return normalize(pts.reshape(-1, 2))


def random_unit(rng):
    """(segments, workspace) of one placement unit; segments overlap the workspace at least once"""
    while True:
        span = int(rng.choice([2000, 20000, 1000000, 50000000]))
        ws = random_workspace(rng, span, int(rng.integers(1, 9)))
        segs = random_list(rng, span, int(rng.integers(1, 80)), int(rng.choice([5, 50, 500, 3000])))
        from gat_b200.segmentlist import SegmentList
        t = SegmentList(array=segs); t._normalized = True
        w = SegmentList(array=ws); w._normalized = True
        t.filter(w)
        if len(t):
            return segs, ws


def random_problem(rng, n_contigs=3, n_iso=0, n_annot=4, span=200000, nseg=60, nanno=80, seglen=300, annolen=1500):
    """a small whole problem: units (optionally split into isochore-like units), contig-level annotations.
    Returns dict with unit_contig, unit_segments, unit_workspace, annotations[a][c], cws_nseg, has_isochores"""
    unit_contig, unit_segments, unit_workspace = [], [], []
    annotations = [[None] * n_contigs for _ in range(n_annot)]
    cws = []
    from gat_b200.segmentlist import SegmentList
    for c in range(n_contigs):
        segs = random_list(rng, span, nseg, seglen)
        if n_iso == 0:
            ws = random_workspace(rng, span, int(rng.integers(1, 4)))
            pieces = [ws]
        else:
            # tile the contig into n_iso interleaved isochore classes
            tile = span // (4 * n_iso)
            bounds = np.arange(0, span + 1, tile)
            labels = rng.integers(0, n_iso, len(bounds) - 1)
            pieces = []
            for i in range(n_iso):
                m = labels == i
                pieces.append(normalize(np.stack([bounds[:-1][m], bounds[1:][m]], axis=1)))
        allws = []
        for ws in pieces:
            if len(ws) == 0:
                continue
            t = SegmentList(array=segs); t._normalized = True
            w = SegmentList(array=ws); w._normalized = True
            t.filter(w)
            allws.append(ws)
            if len(t) == 0:
                continue
            unit_contig.append(c)
            unit_segments.append(t.asarray().copy())
            unit_workspace.append(ws)
        cw = SegmentList(array=np.concatenate(allws)) if allws else SegmentList()
        cw.merge(0) if n_iso else cw.normalize()
        cws.append(len(cw))
        for a in range(n_annot):
            an = SegmentList(array=random_list(rng, span, nanno, annolen)); an._normalized = True
            an.intersect(cw)
            if n_iso:
                an.merge(0)
            annotations[a][c] = an.asarray().copy()
    # contigs are numbered by first appearance among the units
    order = []
    for c in unit_contig:
        if c not in order:
            order.append(c)
    remap = dict((c, i) for i, c in enumerate(order))
    unit_contig = [remap[c] for c in unit_contig]
    annotations = [[annotations[a][c] for c in order] for a in range(n_annot)]
    cws = [cws[c] for c in order]
    return dict(unit_contig=unit_contig, unit_segments=unit_segments, unit_workspace=unit_workspace,
                annotations=annotations, cws_nseg=cws, has_isochores=n_iso > 0, n_contigs=len(order))


def new_context_like(ctx):
    """a fresh context of the same kind as `ctx` (the library reads its environment knobs when a context is created):
    cuda:0, or the SIMT-emulated build when `ctx` is the emulated one (tests/test_emu_parity.py)"""
    if hasattr(ctx.lib, "gatb_emulation_marker"):
        import emu_context
        return emu_context.context()
    from gat_b200 import device
    return device.Context(0)
