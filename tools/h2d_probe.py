import torch, time
x = torch.empty(159*1024*1024, dtype=torch.uint8).pin_memory()
d = torch.empty_like(x, device='cuda')
for _ in range(3):
    d.copy_(x, non_blocking=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    d.copy_(x, non_blocking=True)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print("H2D 159 MiB pinned: %.2f ms = %.1f GB/s" % (ms, 159*1.048576/ms))
h = torch.empty(16*1024*1024, dtype=torch.uint8).pin_memory()
dd = torch.empty_like(h, device='cuda')
e0.record()
for _ in range(10):
    h.copy_(dd, non_blocking=True)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print("D2H 16 MiB pinned: %.2f ms = %.1f GB/s" % (ms, 16*1.048576/ms))
