"""K5 on a large device-resident matrix: the TMA-streamed passes (stats_stream.cu) against the column-tiled kernels
(count.cu), same inputs, results compared; kernel times from gatb_profile (CUDA events around every launch).

    python tools/stats_bench.py [--samples 1000000] [--cols 1000]        # the north-star matrix: 4 GB
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=1000000)
    ap.add_argument("--cols", type=int, default=1000)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    from gat_b200 import device
    S, A = args.samples, args.cols
    g = torch.Generator(device="cuda").manual_seed(5)
    # counts like the benchmark's: a few million bases of overlap, different scale per column
    scale = torch.randint(1000, 4000000, (1, A), device="cuda", generator=g, dtype=torch.int64)
    m = (torch.rand((S, A), device="cuda", generator=g) * scale).to(torch.int64).to(torch.uint32)
    observed = m[0].cpu().numpy().astype(np.float64)
    out = {"samples": S, "cols": A, "matrix_bytes": S * A * 4}
    res = {}
    for name, env in (("streamed", "1"), ("column_tiled", "0")):
        os.environ["GATB_STATS_STREAM"] = env
        ctx = device.Context(0)
        res[name] = ctx.column_stats(None, observed, device_ptr=m.data_ptr(), n_samples=S, n_cols=A, is_float=False)
        ctx.profile(True)
        ctx.profile_read()
        for _ in range(args.reps):
            ctx.column_stats(None, observed, device_ptr=m.data_ptr(), n_samples=S, n_cols=A, is_float=False)
        ms, launches = ctx.profile_read()["other"]
        ctx.profile(False)
        ms, launches = ms / args.reps, launches // args.reps
        # streamed: every profiled scope (pass 1, each select pass) reads the matrix once; tiled: pass 1, pass 2 and the
        # four radix passes inside its select kernel
        passes = launches if name == "streamed" else 6
        out[name] = {"kernel_ms": ms, "launches": launches, "passes_over_matrix": passes,
                     "GBps": passes * S * A * 4.0 / (ms / 1000.0) / 1e9}
        ctx.close()
    same = {}
    for key in ("expected", "lower95", "upper95", "fold", "pvalue"):
        same[key] = bool(np.array_equal(res["streamed"][key], res["column_tiled"][key]))
    same["stddev_max_rel_diff"] = float(np.max(np.abs(res["streamed"]["stddev"] - res["column_tiled"]["stddev"]) /
                                               np.maximum(res["column_tiled"]["stddev"], 1e-300)))
    out["identical"] = same
    out["speedup"] = out["column_tiled"]["kernel_ms"] / out["streamed"]["kernel_ms"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
