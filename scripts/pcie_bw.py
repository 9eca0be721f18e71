import torch, time
n = 829444329
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h = torch.empty(n, dtype=torch.uint8).pin_memory()
for _ in range(2): h.copy_(d, non_blocking=True); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5): h.copy_(d, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 5
print("D2H pinned %.1f MB in %.2f ms = %.1f GB/s" % (n / 1e6, dt * 1e3, n / dt / 1e9))
t0 = time.perf_counter()
for _ in range(5): d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 5
print("H2D pinned %.2f ms = %.1f GB/s" % (dt * 1e3, n / dt / 1e9))
