"""Time the counts-table formatter (gatb_format_counts) at the C5 shape: 1e6 samples x 200 annotations,
next to the reference's Python expression on one column -- measurement aid

    python tools/format_bench.py [n_samples] [n_cols]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402
import torch  # noqa: E402
from gat_b200 import device  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
A = int(sys.argv[2]) if len(sys.argv) > 2 else 200
ctx = device.Context(0)
g = torch.Generator(device="cuda").manual_seed(1)
m = (torch.randn((S, A), device="cuda", generator=g) * 3000 + 26000).clamp_(0).to(torch.int32)
torch.cuda.synchronize()
for it in range(2):
    t0 = time.perf_counter()
    text, off = ctx.format_counts(device_ptr=m.data_ptr(), n_samples=S, n_cols=A)
    dt = time.perf_counter() - t0
print("gatb_format_counts: %i x %i -> %.1f MB of text in %.3f s (%.2e numbers/s, incl. the copy to the host)"
      % (S, A, text.size / 1e6, dt, S * A / dt))
col = m[:, 0].cpu().numpy()
t0 = time.perf_counter()
ref = ",".join(["%i" % x for x in col])
dt1 = time.perf_counter() - t0
assert ref == text[int(off[0]):int(off[1])].tobytes().decode("ascii")
print("reference expression on one column of %i: %.3f s (%.2e numbers/s on one core) -> %.0f s for the table"
      % (S, dt1, S / dt1, dt1 * A))
ctx.close()
