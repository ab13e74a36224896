"""readers of the golden fixtures written by tests/golden/make_golden.py"""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_json(name):
    with open(os.path.join(GOLDEN, name + ".json")) as f:
        return json.load(f)


def load_npz(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    return z, meta


def undelta(arr):
    """inverse of make_golden.delta: [start - previous start, length] -> [start, end)"""
    a = np.asarray(arr, dtype=np.int64).reshape(-1, 2)
    out = np.empty_like(a)
    out[:, 0] = np.cumsum(a[:, 0])
    out[:, 1] = out[:, 0] + a[:, 1]
    return out.astype(np.uint32)


def collection(z, prefix, names, coded=False):
    """-> gat_b200 IntervalCollection from arrays stored under prefix/track/key"""
    from gat_b200 import engine as Engine
    from gat_b200.segmentlist import SegmentList
    coll = Engine.IntervalCollection(prefix.split("/")[-1])
    for track, key in names:
        arr = z["%s/%s/%s" % (prefix, track, key)]
        s = SegmentList(array=undelta(arr) if coded else arr)
        s._normalized = True
        coll.add(track, key, s)
    return coll


def dictionary(z, prefix, keys, coded=False):
    from gat_b200 import engine as Engine
    from gat_b200.segmentlist import SegmentList
    d = Engine.IntervalDictionary()
    for key in keys:
        arr = z["%s/%s" % (prefix, key)]
        s = SegmentList(array=undelta(arr) if coded else arr)
        s._normalized = True
        d.add(key, s)
    return d
