"""Write the synthetic hg19-shaped inputs of a BASELINE configuration as BED files (column 4 = track), for the
command line path (`python -m gat_b200.cli`) and for the reference's gat-run.py.

    python tools/make_bed.py OUTDIR [--segments 10000] [--annotations 1000] [--isochores]
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402


def write(collection, filename):
    """one pass per collection through Arrow's multi-threaded CSV writer (a Python loop over 2e7 lines takes ~15 s)"""
    import pyarrow as pa
    import pyarrow.csv as pacsv
    contigs, starts, ends, tracks = [], [], [], []
    for track, vv in collection.items():
        for contig, s in vv.items():
            a = s.asarray()
            if len(a) == 0:
                continue
            contigs.append(np.full(len(a), contig))
            tracks.append(np.full(len(a), track))
            starts.append(a[:, 0].astype(np.int64))
            ends.append(a[:, 1].astype(np.int64))
    table = pa.table({"contig": pa.array(np.concatenate(contigs)).dictionary_encode(),
                      "start": np.concatenate(starts), "end": np.concatenate(ends),
                      "track": pa.array(np.concatenate(tracks)).dictionary_encode()})
    pacsv.write_csv(table, filename, write_options=pacsv.WriteOptions(include_header=False, delimiter="\t",
                                                                      quoting_style="none"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("outdir")
    ap.add_argument("--segments", type=int, default=10000)
    ap.add_argument("--annotations", type=int, default=1000)
    ap.add_argument("--annotation-intervals", type=int, default=20000)
    ap.add_argument("--isochores", action="store_true")
    args = ap.parse_args()
    from gat_b200 import synthetic
    t0 = time.time()
    segments, annotations, workspaces, iso = synthetic.make(args.segments, args.annotations, args.annotation_intervals,
                                                            isochores=args.isochores)
    os.makedirs(args.outdir, exist_ok=True)
    write(segments, os.path.join(args.outdir, "segments.bed"))
    write(annotations, os.path.join(args.outdir, "annotations.bed"))
    write(workspaces, os.path.join(args.outdir, "workspace.bed"))
    if iso is not None:
        write(iso, os.path.join(args.outdir, "isochores.bed"))
    sys.stderr.write("wrote %s in %.1f s\n" % (args.outdir, time.time() - t0))


if __name__ == "__main__":
    main()
