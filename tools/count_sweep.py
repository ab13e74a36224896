#!/usr/bin/env python
"""Sweep the count kernel's knobs on the bench workload in ONE process (the workload is built once).

    python tools/count_sweep.py "" "SHIFT=9" "THREADS=512,SCHUNK=14" ...

Knobs (environment variables read by the library at context / annotation-set creation):
KGRP -> GATB_KEY_GROUP, SHIFT -> GATB_BIN_SHIFT, SCHUNK -> GATB_SCHUNK,
THREADS -> GATB_COUNT_THREADS, GROUP -> GATB_GROUP_TRACKS.  Prints per configuration the CUDA-event time of
the count and placement kernels (3 profiled steps after 2 warm-up steps) and the device-resident step time.
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

ENV = {"KGRP": "GATB_KEY_GROUP", "SHIFT": "GATB_BIN_SHIFT", "SCHUNK": "GATB_SCHUNK",
       "THREADS": "GATB_COUNT_THREADS", "GROUP": "GATB_GROUP_TRACKS", "BATCH": "BATCH"}


def main():
    import torch
    import bench
    from gat_b200 import device
    sys.argv = [sys.argv[0]] + [a for a in sys.argv[1:] if a.startswith("--")]
    configs = [a for a in sys.argv_orig if not a.startswith("--")]
    args = bench.parse_args()
    wl = bench.build_workload(args)
    pr, A, C = wl["problem"], wl["A"], wl["C"]
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    for cfg in configs or [""]:
        for v in ENV.values():
            os.environ.pop(v, None)
        B = args.samples_per_step
        for kv in filter(None, cfg.split(",")):
            k, v = kv.split("=")
            if k == "BATCH":
                B = int(v)
            else:
                os.environ[ENV[k]] = v
        ctx = device.Context(0)
        ctx.set_stream(stream.cuda_stream)
        ctx.set_batch_size(B)
        t0 = time.perf_counter()
        annos = device.Annotations(ctx, None, key_ws_nseg=wl["nseg"], csr=(A, C) + wl["anno_csr"])
        t_an = time.perf_counter() - t0
        smp = device.Sampler(ctx, pr.unit_contig, C, pr.has_isochores, None, None, csr=(wl["seg_csr"], wl["ws_csr"]))
        out_u = torch.zeros((1, B, A), dtype=torch.int32, device=dev)

        def step(i):
            return smp.run(annos, [args.counter], 20260101, 0, i * B, B, out_counts_ptr=out_u.data_ptr())

        for i in range(2):
            step(i)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 5
        e0.record(stream)
        for i in range(n):
            step(2 + i)
        e1.record(stream)
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / n
        ctx.profile(True)
        for i in range(3):
            step(10 + i)
        prof = ctx.profile_read()
        ctx.profile(False)
        chk = int(out_u.to(torch.int64).sum().item())
        print("%-40s count %7.3f ms  place %6.3f ms  step %7.3f ms  %9.0f samples/s  annos %.0f ms  checksum %d"
              % (cfg or "(defaults)", prof["count"][0] / max(prof["count"][1], 1),
                 prof["place"][0] / max(prof["place"][1], 1), ms, B / ms * 1e3, t_an * 1e3, chk), flush=True)
        smp.close()
        annos.close()
        ctx.close()


if __name__ == "__main__":
    sys.argv_orig = sys.argv[1:]
    main()
