"""PCIe copy bandwidth between pinned host memory and the GPU (the ceilings of the host-pointer path)."""
import time, torch
n = 1 << 28                                   # 256 MiB
h = torch.empty(n, dtype=torch.uint8).pin_memory(); h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda"); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=8):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return reps * n / (time.perf_counter() - t0) / 1e9
print(f"H2D {t(lambda: d.copy_(h, non_blocking=True)):.1f} GB/s")
print(f"D2H {t(lambda: h.copy_(d, non_blocking=True)):.1f} GB/s")
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
print(f"H2D + D2H concurrently: {t(both):.1f} GB/s each direction")
for kb in (64, 512, 4096):
    m = kb << 10
    def small():
        for i in range(0, 1 << 26, m): h[i:i + m].copy_(d[i:i + m], non_blocking=True)
    small(); torch.cuda.synchronize(); t0 = time.perf_counter(); small(); torch.cuda.synchronize()
    print(f"D2H in {kb} KiB pieces: {(1 << 26) / (time.perf_counter() - t0) / 1e9:.1f} GB/s")
