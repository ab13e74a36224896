"""Device-side input preparation (gat_b200/csrc/prep.cu, Engine.DeviceIntervalCollection) against
the host implementation of the same reference operations (gat_b200/segmentlist.py, pinned to the reference by
tests/test_oracle_golden.py / tests/test_host_logic.py), against the CPU oracle, and against the lists the
reference's own IO produced (tests/golden/prep_isochores.json).  Index work: bit-exact."""
import numpy as np
import pytest

from tests import golden_util as G

pytestmark = pytest.mark.gpu


def _rows(rng, n_lists, n_rows, span, maxlen, empties=True):
    l = rng.integers(0, n_lists, n_rows)
    s = rng.integers(0, span, n_rows)
    length = rng.integers(0 if empties else 1, maxlen + 1, n_rows)
    if n_rows > 10:                                     # exact duplicates, adjacency, containment
        s[1], l[1], length[1] = s[0], l[0], length[0]
        s[3], l[3] = s[2] + length[2], l[2]
        s[5], l[5], length[5] = s[4] + 1, l[4], max(0, int(length[4]) - 2)
    return l.astype(np.uint32), s.astype(np.uint32), (s + length).astype(np.uint32)


def _split(lists):
    offs, data = lists.download()
    return [data[int(offs[i]):int(offs[i + 1])] for i in range(lists.n_lists)]


def test_lists_from_rows_normalize_and_merge(ctx, oracle):
    """rows in file order -> SegmentList.normalize / merge(0) of every list (gat/SegmentList.pyx:697-816)"""
    from gat_b200 import device
    rng = np.random.default_rng(1)
    for n_lists, n_rows, span, maxlen in ((1, 40, 300, 30), (7, 3000, 20000, 60), (300, 200000, 5000000, 4000),
                                          (5, 0, 10, 3), (4, 1, 10, 3)):
        l, s, e = _rows(rng, n_lists, n_rows, span, maxlen)
        for join in (False, True):
            got = device.Lists.from_rows(ctx, l, s, e, n_lists, join_adjacent=join)
            lists = _split(got)
            count, bases = got.sizes()
            got.close()
            for i in range(n_lists):
                raw = np.stack([s[l == i], e[l == i]], axis=1)
                want = oracle.merge(raw, 0) if join else oracle.normalize(raw)
                assert np.array_equal(lists[i], want), (n_lists, n_rows, join, i)
                assert int(count[i]) == len(want) and int(bases[i]) == int((want[:, 1].astype(np.int64) - want[:, 0]).sum())


def test_lists_restrict_collapse_select(ctx, oracle):
    """intersect / filter against the workspace (fanout 1) and against isochore tracks (fanout F), fromIsochores,
    and key re-ordering -- each list against the oracle's SegmentList operations"""
    from gat_b200 import device
    rng = np.random.default_rng(2)
    for T, K, F in ((3, 4, 1), (5, 3, 4), (1, 1, 2)):
        n_in = T * K
        l, s, e = _rows(rng, n_in, 400 * n_in, 300000, 900, empties=False)
        lists = device.Lists.from_rows(ctx, l, s, e, n_in)
        mine = _split(lists)
        others = []
        for k in range(K):
            if F == 1:
                pts = np.sort(rng.choice(300000, size=2 * int(rng.integers(0, 9)), replace=False))
                others.append(oracle.normalize(pts.reshape(-1, 2)))
            else:                                   # tiles with random labels: the isochore tracks partition the key
                bounds = np.arange(0, 300001, 5000)
                label = rng.integers(0, F, len(bounds) - 1)
                for f in range(F):
                    m = label == f
                    others.append(oracle.normalize(np.stack([bounds[:-1][m], bounds[1:][m]], axis=1)))
        other = device.Lists.from_lists(ctx, others)
        for truncate in (True, False):
            out = lists.restrict(K, F, other, truncate)
            got = _split(out)
            assert out.n_lists == n_in * F
            for i in range(n_in):
                for f in range(F):
                    o = others[(i % K) * F + f]
                    want = oracle.intersect(mine[i], o) if truncate else oracle.filter_(mine[i], o)
                    assert np.array_equal(got[i * F + f], want), (T, K, F, truncate, i, f)
            if F > 1:
                back = out.collapse(F)              # fromIsochores: extend in isochore order, merge(0)
                got_back = _split(back)
                for i in range(n_in):
                    cat = np.concatenate([got[i * F + f] for f in range(F)]) if F else mine[i]
                    assert np.array_equal(got_back[i], oracle.merge(cat, 0)), (truncate, i)
                back.close()
            out.close()
        src = rng.permutation(n_in + 2).astype(np.uint32)       # two indices beyond the end: empty lists
        sel = lists.select(src)
        got = _split(sel)
        for j, i in enumerate(src):
            assert np.array_equal(got[j], mine[i] if i < n_in else np.zeros((0, 2), dtype=np.uint32))
        sel.close()
        other.close()
        lists.close()


def test_invalid_rows_fail_loudly(ctx):
    from gat_b200 import device, _lib
    z = np.zeros(1, dtype=np.uint32)
    with pytest.raises(_lib.GatB200Error) as e:
        device.Lists.from_rows(ctx, z, np.array([5], dtype=np.uint32), np.array([2], dtype=np.uint32), 1)
    assert e.value.code == _lib.ERR_INVALID
    with pytest.raises(_lib.GatB200Error) as e:
        device.Lists.from_rows(ctx, z, z, np.array([2 ** 31], dtype=np.uint32), 1)
    assert e.value.code == _lib.ERR_RANGE
    with pytest.raises(_lib.GatB200Error):
        device.Lists.from_rows(ctx, np.array([3], dtype=np.uint32), z, np.array([2], dtype=np.uint32), 3)


def test_device_preparation_matches_reference_io(ctx, tmp_path):
    """IO.buildSegments + IO.applyIsochores with the annotations prepared ON THE GPU give the lists the reference's
    own IO gave for the same BED files, isochore file included (tests/golden/prep_isochores.json, both cases)"""
    import gat_b200
    from gat_b200 import io as IO, engine
    data = G.load_json("prep_isochores")
    argv = []
    for name, flag in (("segments", "--segments"), ("annotations", "--annotations"), ("workspace", "--workspace"),
                       ("iso", "--isochore-file")):
        path = tmp_path / (name + ".bed")
        path.write_text(data["files"][name + ".bed"])
        argv.append("%s=%s" % (flag, path))
    for case in data["cases"]:
        options, _ = gat_b200.buildParser().parse_args(argv + case["extra"])
        segments, annotations, workspaces, isochores = IO.buildSegments(options)
        assert isinstance(annotations, engine.DeviceIntervalCollection) and annotations.onDevice
        workspace = IO.applyIsochores(segments, annotations, workspaces, options, isochores,
                                      truncate_segments_to_workspace=options.truncate_segments_to_workspace)
        assert annotations.onDevice
        assert sorted(annotations.tracks) == sorted(case["annotations"].keys())
        total = 0
        for track, lists in case["annotations"].items():
            for k, want in lists.items():
                got = annotations[track][k].asList() if k in annotations[track] else []
                assert got == [tuple(x) for x in want], (track, k)
                total += sum(e - s for s, e in want)
        assert annotations.sum() == total
        sizes = annotations.trackSizes()
        for track, lists in case["annotations"].items():
            assert sizes[track] == (sum(len(v) for v in lists.values()),
                                    sum(e - s for v in lists.values() for s, e in v))
        # fromIsochores of a clone (what the sampling loop counts against): the host result of the same lists
        contig = annotations.clone()
        contig.fromIsochores()
        host = engine.IntervalCollection()
        for track, lists in case["annotations"].items():
            for k, want in lists.items():
                host.add(track, k, gat_b200.SegmentList(array=np.array(want, dtype=np.uint32).reshape(-1, 2)))
        host.fromIsochores()
        for track in host.tracks:
            for k, want in host[track].items():
                got = contig[track][k].asList() if k in contig[track] else []
                assert got == want.asList(), (track, k)


@pytest.mark.parametrize("with_isochores", [False, True])
def test_cli_device_and_host_preparation_agree(ctx, tmp_path, monkeypatch, with_isochores):
    """the whole command line run with the annotations prepared on the GPU prints the table of the run with host
    preparation, byte for byte (observed counts, sampled statistics, size and overlap columns)"""
    from gat_b200 import cli, synthetic
    genome = [("chrA", 1500000), ("chrB", 700000), ("chrC", 90000)]
    segments, annotations, workspaces, iso = synthetic.make(
        n_segments=400, n_annotations=7, n_annotation_intervals=500, isochores=with_isochores, genome=genome,
        isochore_tile=50000, n_isochores=3, seed=5)
    files = {}
    for name, coll in (("segments", segments), ("annotations", annotations), ("workspace", workspaces)):
        files[name] = str(tmp_path / (name + ".bed"))
        synthetic.write_bed(coll, files[name], with_tracks=name == "annotations")
    argv = ["gat-run", "--segments=" + files["segments"], "--annotations=" + files["annotations"],
            "--workspace=" + files["workspace"], "--counter=nucleotide-overlap", "--counter=nucleotide-density",
            "--counter=segment-overlap", "--num-samples=200", "--random-seed=3", "-v", "0"]
    if with_isochores:
        files["iso"] = str(tmp_path / "iso.bed")
        synthetic.write_bed(iso, files["iso"], with_tracks=True)
        argv.append("--isochore-file=" + files["iso"])
    tables = {}
    for mode in ("device", "host"):
        if mode == "host":
            monkeypatch.setenv("GATB_HOST_PREP", "1")
        pattern = str(tmp_path / (mode + "_%s.tsv"))
        assert cli.main(argv + ["--output-tables-pattern=" + pattern]) == 0
        tables[mode] = [open(pattern % c).read() for c in ("nucleotide-overlap", "nucleotide-density", "segment-overlap")]
    assert tables["device"] == tables["host"]
    assert tables["device"][0].count("\n") == 8
