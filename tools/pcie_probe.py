"""Raw pinned D2H / H2D bandwidth of the box (the ceiling of the host-buffer walk API)."""
import torch, time
n = 1337 << 20
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h = torch.empty(n, dtype=torch.uint8).pin_memory()
for name, fn in (("D2H", lambda: h.copy_(d, non_blocking=True)), ("H2D", lambda: d.copy_(h, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print("%s %.1f MB in %.2f ms = %.1f GB/s" % (name, n / 1e6, dt * 1e3, n / dt / 1e9))
# chunked D2H (48 MB pieces) like gw_node2vec_walks
c = 48 << 20
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5):
    for lo in range(0, n, c):
        h[lo:lo + c].copy_(d[lo:lo + c], non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 5
print("D2H in 48 MB chunks: %.2f ms = %.1f GB/s" % (dt * 1e3, n / dt / 1e9))
