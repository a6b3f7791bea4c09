"""PCIe / pinned-copy diagnostic: H2D, D2H, and both concurrently (GB/s)."""
import torch, time
n = 1 << 30
h1 = torch.empty(n, dtype=torch.uint8).pin_memory(); h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d1 = torch.empty(n, dtype=torch.uint8, device="cuda"); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=5):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
a = t(lambda: d1.copy_(h1, non_blocking=True)); print(f"H2D {n/a/1e9:.1f} GB/s")
b = t(lambda: h2.copy_(d2, non_blocking=True)); print(f"D2H {n/b/1e9:.1f} GB/s")
def both():
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
c = t(both); print(f"H2D+D2H concurrent: {2*n/c/1e9:.1f} GB/s total ({n/c/1e9:.1f} each)")
