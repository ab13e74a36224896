"""Parity at the shapes that are MEASURED: the bench / north-star shape (10k segments x 1000
tracks, one 4096-sample batch, samples at the CTA-chunk edges of the counting kernel), BASELINE config 4 (50k
segments x 1000 tracks), config 5 (global sample indices around 10^6, 200 tracks) and config 3 at its full 50
tracks (8 GC isochores) -- every one against the sequential CPU oracle under the same Philox stream, bit-exact.

The oracle costs ~0.5-1.5 s per sample at these sizes; building the 1000-track synthetic input ~10 s.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ALL = ["nucleotide-overlap", "nucleotide-density", "segment-overlap", "segment-midoverlap",
       "annotation-overlap", "annotation-midoverlap"]


@pytest.fixture(scope="module")
def northstar(ctx):
    """bench.py's workload: synthetic hg19, 10 000 segments x 1 000 annotation tracks of 20 000 intervals"""
    import gat_b200
    from gat_b200 import synthetic, device
    segments, annotations, workspaces, _ = synthetic.make(10000, 1000, 20000)
    workspace = synthetic.prepare(segments, annotations, workspaces)
    pr = gat_b200.TrackProblem(segments["merged"], workspace)
    atracks, lists, nseg = gat_b200.buildContigAnnotations(annotations, workspace, pr.contigs)
    assert len(atracks) == 1000 and len(pr.contigs) == 24
    smp = device.Sampler(ctx, pr.unit_contig, len(pr.contigs), False, pr.unit_segments, pr.unit_workspace)
    annos = device.Annotations(ctx, lists, key_ws_nseg=nseg)
    yield dict(problem=pr, lists=lists, nseg=nseg, smp=smp, annos=annos, contigs=pr.contigs)
    smp.close()
    annos.close()


def _check(oracle, pr, lists, nseg, names, res, seed, track, begin, picks, has_isochores=False):
    for s in picks:
        exp = oracle.compute_sample_philox(pr.unit_contig, pr.unit_segments, pr.unit_workspace, lists, nseg, names,
                                           seed=seed, track=track, sample=begin + s, has_isochores=has_isochores)
        for i, n in enumerate(names):
            got = np.asarray(res[n][s], dtype=np.float64)
            assert np.array_equal(got, exp[i]), "sample %i (global %i), counter %s: %i of %i columns differ" % (
                s, begin + s, n, int((got != exp[i]).sum()), len(got))


def test_bench_shape_batch_matches_oracle(ctx, northstar, oracle):
    """the bench step itself: ONE gatb_run of 4096 samples x 1000 tracks (ka = 1000, one track group, 28 samples
    per counting CTA), all six counters; samples at the first / last slot of the first, a middle and the last
    CTA chunk against the oracle"""
    pr, smp, annos = northstar["problem"], northstar["smp"], northstar["annos"]
    B = 4096
    ctx.set_batch_size(B)
    res, info = smp.run(annos, ALL, seed=20260101, track=0, sample_begin=0, n_samples=B)
    ctx.set_batch_size(0)
    assert int(info[2]) == 0
    chunk = -(-B // 148)                    # samples per counting CTA (count_params_annos)
    last = (B - 1) // chunk * chunk
    picks = [0, chunk - 1, chunk, 49 * chunk - 1, 73 * chunk, last - 1, last, B - 1]
    _check(oracle, pr, northstar["lists"], northstar["nseg"], ALL, res, 20260101, 0, 0, picks)
    # the statistic the bench reports next to the rate
    assert 9000 * B < int(info[0]) < 11000 * B


def test_bench_shape_multi_batch_and_offset(ctx, northstar, oracle):
    """the same shape the way bench.py --gpus N runs it: sample_begin = (step * world + rank) * B, several
    internal batches per call (a step of 3 x 4096 samples); the rows equal single-sample runs of the same
    global indices"""
    pr, smp, annos = northstar["problem"], northstar["smp"], northstar["annos"]
    names = ["nucleotide-overlap"]
    begin = (7 * 8 + 5) * 4096
    ctx.set_batch_size(4096)
    res, _ = smp.run(annos, names, seed=20260101, track=0, sample_begin=begin, n_samples=3 * 4096)
    ctx.set_batch_size(0)
    _check(oracle, pr, northstar["lists"], northstar["nseg"], names, res, 20260101, 0, begin,
           [0, 4095, 4096, 8191, 8192, 3 * 4096 - 1])


def test_config5_sample_indices_near_1e6(ctx, northstar, oracle):
    """BASELINE config 5: 10k segments x 200 tracks, 10^6 samples -- the last ten global sample indices"""
    from gat_b200 import device
    pr, smp = northstar["problem"], northstar["smp"]
    lists = northstar["lists"][:200]
    annos = device.Annotations(ctx, lists, key_ws_nseg=northstar["nseg"])
    names = ["nucleotide-overlap", "segment-overlap"]
    res, info = smp.run(annos, names, seed=1, track=0, sample_begin=999990, n_samples=10)
    annos.close()
    assert int(info[2]) == 0
    _check(oracle, pr, lists, northstar["nseg"], names, res, 1, 0, 999990, range(10))


def test_config4_50k_segments_1000_tracks(ctx, northstar, oracle):
    """BASELINE config 4 (ENCODE scale): 50 000 segments x 1000 tracks; two samples of a 64-sample run"""
    import gat_b200
    from gat_b200 import synthetic, device, engine
    segments, _, workspaces, _ = synthetic.make(50000, 0, 0)
    workspace = synthetic.prepare(segments, engine.IntervalCollection("annotations"), workspaces)
    pr = gat_b200.TrackProblem(segments["merged"], workspace)
    assert list(pr.contigs) == list(northstar["contigs"])
    assert sum(len(x) for x in pr.unit_segments) > 49000
    smp = device.Sampler(ctx, pr.unit_contig, len(pr.contigs), False, pr.unit_segments, pr.unit_workspace)
    names = ["nucleotide-overlap", "segment-overlap", "annotation-overlap"]
    res, info = smp.run(northstar["annos"], names, seed=4, track=0, sample_begin=12500, n_samples=64)
    smp.close()
    assert int(info[2]) == 0 and int(info[0]) > 64 * 45000
    _check(oracle, pr, northstar["lists"], northstar["nseg"], names, res, 4, 0, 12500, [0, 63])


def test_config3_isochores_50_tracks(ctx, oracle):
    """BASELINE config 3 at its full size: 10k segments, 50 tracks, 8 GC isochores (<= 192 units), the config's
    segment-overlap counter plus nucleotide-overlap and the float64 density; 8 samples of a 300-sample run"""
    import gat_b200
    from gat_b200 import synthetic, device
    segments, annotations, workspaces, iso = synthetic.make(10000, 50, 20000, isochores=True)
    workspace = synthetic.prepare(segments, annotations, workspaces, iso)
    pr = gat_b200.TrackProblem(segments["merged"], workspace)
    assert pr.has_isochores and len(pr.contigs) == 24
    atracks, lists, nseg = gat_b200.buildContigAnnotations(annotations, workspace, pr.contigs)
    assert len(atracks) == 50
    smp = device.Sampler(ctx, pr.unit_contig, len(pr.contigs), True, pr.unit_segments, pr.unit_workspace)
    annos = device.Annotations(ctx, lists, key_ws_nseg=nseg)
    names = ["segment-overlap", "nucleotide-overlap", "nucleotide-density"]
    res, info = smp.run(annos, names, seed=3, track=0, sample_begin=0, n_samples=300)
    smp.close()
    annos.close()
    assert int(info[2]) == 0
    _check(oracle, pr, lists, nseg, names, res, 3, 0, 0, [0, 1, 2, 147, 148, 149, 298, 299], has_isochores=True)
