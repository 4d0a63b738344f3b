"""Per-tile fixed cost of ozaki_mma_kernel: short-k products (few MMA steps per tile) at several digit counts."""
import sys
import torch
sys.path.insert(0, ".")
from gaussian_process_optimization_b200 import native  # noqa: E402


def ev_time(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


n = 8192
g = torch.Generator(device="cuda").manual_seed(0)
C = torch.zeros(n, n, dtype=torch.float64, device="cuda")
for k in (128, 256, 1024):
    A = torch.randn(n, k, dtype=torch.float64, device="cuda", generator=g)
    for S in (1, 2, 4, 7):
        t = ev_time(lambda: native.ozaki_dgemm(0, 0, 1.0, A, A, 0.0, C, slices=S))
        steps = S * (S + 1) // 2 * (k // 128)
        tiles_per_sm = 1024 / 148
        print("k=%d S=%d: %.3f ms total; per tile %.1f us for %d MMA steps (%.1f us at 0.563 us/step) and %d drains"
              % (k, S, t, t * 1e3 / tiles_per_sm, steps, steps * 0.563, S), flush=True)
