"""Throughput of --sampler=shift on the benchmark shape (10k segments x 1000 annotation tracks, contig
workspace): CUDA-event time of shift_kernel and of a whole step (placement + counting), per 4096 samples.

    python tools/shift_bench.py [--annotations 1000] [--isochores]
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import bench  # noqa: E402
from gat_b200 import device  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--annotations", type=int, default=1000)
ap.add_argument("--isochores", action="store_true")
ap.add_argument("--radius", type=float, default=2.0)
a = ap.parse_args()
args = argparse.Namespace(segments=10000, annotations=a.annotations, annotation_intervals=20000,
                          isochores=a.isochores, counter="nucleotide-overlap", samples_per_step=4096)
wl = bench.build_workload(args)
pr, A, C, B = wl["problem"], wl["A"], wl["C"], 4096
ctx = device.Context(0)
ctx.set_batch_size(B)
annos = device.Annotations(ctx, None, key_ws_nseg=wl["nseg"], csr=(A, C) + wl["anno_csr"])
smp = device.Sampler(ctx, pr.unit_contig, C, pr.has_isochores, None, None, csr=(wl["seg_csr"], wl["ws_csr"]), bucket_size=0)
smp.set_shift(a.radius, 0)
import torch  # noqa: E402
out = torch.zeros((1, B, A), dtype=torch.int32, device="cuda:0")
for i in range(3):
    info = smp.run(annos, ["nucleotide-overlap"], 1, 0, i * B, B, out_counts_ptr=out.data_ptr())
ctx.profile(True)
torch.cuda.synchronize()
t0 = time.perf_counter()
n = 20
for i in range(n):
    info = smp.run(annos, ["nucleotide-overlap"], 1, 0, (3 + i) * B, B, out_counts_ptr=out.data_ptr())
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / n
prof = ctx.profile_read()
print("sampler=shift radius=%g %s: shift_kernel %.3f ms, merge %.3f ms, count %.3f ms per %i samples; step %.3f ms = %.0f samples/s; "
      "pieces per sample %.1f; checksum %i"
      % (a.radius, "isochores" if a.isochores else "contigs", prof["place"][0] / max(prof["place"][1], 1),
         prof["merge"][0] / max(prof["merge"][1], 1), prof["count"][0] / max(prof["count"][1], 1), B, dt * 1e3,
         B / dt, float(info[0]) / B, int(out.sum().item())))
