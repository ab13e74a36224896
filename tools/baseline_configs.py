"""Run the BASELINE.json configurations end to end through the public API (gat_b200.run + outputResults)
and print one JSON line per configuration (wall seconds, samples/s).  Profiling aid, not a test.

    python tools/baseline_configs.py [c2 c3 c4 c5]
"""
import hashlib
import io
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import gat_b200  # noqa: E402
from gat_b200 import engine as Engine, io as IO, synthetic, parallel  # noqa: E402

CONFIGS = {
    # name: (segments, annotations, isochores, counter, samples)
    "c2": (10000, 50, False, "nucleotide-overlap", 10000),
    "c3": (10000, 50, True, "segment-overlap", 10000),
    "c4": (50000, 1000, False, "nucleotide-overlap", 12500),      # one GPU's share of 100k samples on 8 GPUs
    "c4full": (50000, 1000, False, "nucleotide-overlap", 100000),  # BASELINE config 4 as stated (meant for 8 GPUs)
    "c5": (10000, 200, False, "nucleotide-overlap", 1000000),
    "ns": (10000, 1000, False, "nucleotide-overlap", 100000),     # north-star shape, 1e5 samples
}


class Opt(object):
    qvalue_method = "BH"
    qvalue_lambda = None
    qvalue_pi0_method = "smoother"
    output_order = "fold"
    output_tables_pattern = "/tmp/gat_b200_%s.tsv"


def main():
    parallel.init_from_env()
    rank, world = parallel.rank_world()
    names = sys.argv[1:] or ["c2", "c3", "c5"]
    for name in names:
        nseg, nanno, iso, counter, S = CONFIGS[name]
        t0 = time.perf_counter()
        segments, annotations, workspaces, isochores = synthetic.make(nseg, nanno, 20000, isochores=iso)
        workspace = synthetic.prepare(segments, annotations, workspaces, isochores)
        t_prep = time.perf_counter() - t0
        Engine.seed(1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = gat_b200.run(segments, annotations, workspace, Engine.SamplerAnnotator(bucket_size=1, nbuckets=100000),
                           [Engine.COUNTER_CLASSES[counter]()], Engine.UnconditionalWorkspace(), num_samples=S)
        torch.cuda.synchronize()
        t_run = time.perf_counter() - t0
        t0 = time.perf_counter()
        Opt.stdout = io.StringIO()
        if rank == 0:
            IO.outputResults(res, Opt, Engine.AnnotatorResultExtended.headers, [], 0, {})
        t_out = time.perf_counter() - t0
        if rank == 0:
            rows = Opt.stdout.getvalue().strip().split("\n")
            print(json.dumps({"config": name, "segments": nseg, "annotations": nanno, "isochores": iso, "counter": counter,
                              "samples": S, "gpus": world, "prep_s": round(t_prep, 2), "run_s": round(t_run, 3),
                              "output_s": round(t_out, 3), "samples_per_s": round(S / t_run, 1),
                              "rows": len(rows) - 1, "table_md5": hashlib.md5(Opt.stdout.getvalue().encode()).hexdigest(),
                              "min_p": min(r.pvalue for r in res),
                              "min_q": min(r.qvalue for r in res), "first_row": rows[1][:120]}))
    parallel.finalize()


if __name__ == "__main__":
    main()
