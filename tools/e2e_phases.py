"""time the phases of one end-to-end step (host buffers in, host counts out) -- profiling aid"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from gat_b200 import device

class A: pass
args = A(); args.segments=10000; args.annotations=1000; args.annotation_intervals=20000; args.isochores=False
args.samples_per_step=4096; args.counter="nucleotide-overlap"
wl = bench.build_workload(args)
pr, Aa, C, B = wl["problem"], wl["A"], wl["C"], args.samples_per_step
ctx = device.Context(0)
ctx.set_batch_size(B)
if len(sys.argv) > 1 and sys.argv[1] == 'torchstream':
    ctx.set_stream(torch.cuda.current_stream(torch.device('cuda', 0)).cuda_stream)
    print('using torch current stream')
def pinned(x):
    signed = {np.dtype(np.uint64): np.int64, np.dtype(np.uint32): np.int32}[x.dtype]
    t = torch.from_numpy(np.ascontiguousarray(x).view(signed)).pin_memory()
    return t, t.numpy().view(x.dtype)
keep = []
def P(csr):
    out = []
    for x in csr:
        t, v = pinned(x); keep.append(t); out.append(v)
    return tuple(out)
pa, ps, pw = P(wl["anno_csr"]), P(wl["seg_csr"]), P(wl["ws_csr"])
host = torch.empty((B, Aa), dtype=torch.int32).pin_memory(); hnp = host.numpy().view(np.uint32)
ids = device.counter_ids([args.counter]); info = np.zeros(3, dtype=np.uint64)
lazy = 'lazy' in sys.argv
ctx.profile(True)
for it in range(4):
    t = [time.perf_counter()]
    if lazy: s2 = device.Sampler(ctx, pr.unit_contig, C, pr.has_isochores, None, None, csr=(ps, pw))
    a2 = device.Annotations(ctx, None, key_ws_nseg=wl["nseg"], csr=(Aa, C) + pa, lazy=lazy)
    if not lazy: ctx.synchronize()
    t.append(time.perf_counter())
    if not lazy: s2 = device.Sampler(ctx, pr.unit_contig, C, pr.has_isochores, None, None, csr=(ps, pw)); ctx.synchronize()
    t.append(time.perf_counter())
    ctx.check(ctx.lib.gatb_run(s2.handle, a2.handle, 1, device._p(ids), 1, 0, it * B, B, device._p(hnp), None, 0, device._p(info))); t.append(time.perf_counter())
    s2.close(); t.append(time.perf_counter())
    a2.close(); t.append(time.perf_counter())
    names = ["annotations_create", "sampler_create", "run", "sampler_destroy", "annotations_destroy"]
    print(it, " ".join("%s=%.1fms" % (n, 1e3 * (t[i + 1] - t[i])) for i, n in enumerate(names)), "total=%.1fms" % (1e3 * (t[-1] - t[0])))
print("kernel classes (ms, launches) over 4 steps:", ctx.profile_read())
