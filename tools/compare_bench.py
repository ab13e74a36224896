"""gat-compare statistics: pairs/s of gatb_compare_stats on the GPU next to the reference's per-pair numpy path
(the reference leg is a numpy restatement of scripts/gat-compare.py:218-241 + AnnotatorResult, i.e. the oracle)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gat_b200 import device
from oracle import oracle

S, A = int(sys.argv[1]) if len(sys.argv) > 1 else 10000, int(sys.argv[2]) if len(sys.argv) > 2 else 200
rng = np.random.default_rng(5)
lam = rng.uniform(5, 500, A)
m = rng.poisson(lam, size=(S, A)).astype(np.float64)
obs = rng.poisson(lam * rng.choice([0.5, 1, 2], A)).astype(np.float64)
ctx = device.Context(0)
base = ctx.column_stats(m.astype(np.uint32), obs, pseudo_count=1.0)
pairs = [(i, j) for i in range(A) for j in range(i + 1, A)]
c1, c2 = np.array([p[0] for p in pairs]), np.array([p[1] for p in pairs])
delta = base["fold"][c2] - base["fold"][c1]
for it in range(3):
    t0 = time.perf_counter()
    got = ctx.compare_stats(m, None, c1, c2, obs[c1], obs[c2], delta, pseudo_count=1.0)
    dt = time.perf_counter() - t0
print("GPU: %i pairs x %i samples in %.3f s = %.0f pairs/s (host matrix in, statistics out)" % (len(pairs), S, dt, len(pairs) / dt))
oracle.build()
n = 60
t0 = time.perf_counter()
for q in range(n):
    i, j = pairs[q * (len(pairs) // n)]
    oracle.compare_pair(obs[i], m[:, i], base["fold"][i], obs[j], m[:, j], base["fold"][j], 1.0)
dt = time.perf_counter() - t0
print("CPU (1 core, numpy + C statistics oracle): %i pairs in %.3f s = %.0f pairs/s" % (n, dt, n / dt))
