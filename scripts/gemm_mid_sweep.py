"""Tile configuration at the middle levels of the recursion (256..2048 cubed, full k-range): 64x64 four-CTA (9), 64x64 three-stage
(2), 32x32 (3), 64x128 (1).  Many back-to-back launches per timing so that launch latency is included the way the recursion sees it."""
import sys

import torch

sys.path.insert(0, ".")
from gaussian_process_optimization_b200 import native  # noqa: E402

cfgs = [int(c) for c in sys.argv[1].split(",")] if len(sys.argv) > 1 else [9, 2, 3, 1]
for n in (1024, 1280, 1536, 2048):
    A = torch.randn(n, n, dtype=torch.float64, device="cuda")
    B = torch.randn(n, n, dtype=torch.float64, device="cuda")
    C = torch.zeros(n, n, dtype=torch.float64, device="cuda")
    reps = 20
    row = []
    for cfg in cfgs:
        native.gemm_config(cfg)
        best = 1e30
        for ta, tb in ((0, 0),):
            for _ in range(3):
                native.dgemm(ta, tb, 1.0, A, B, 0.0, C)
            torch.cuda.synchronize()
            for _ in range(3):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(reps):
                    native.dgemm(ta, tb, 1.0, A, B, 0.0, C)
                b.record()
                torch.cuda.synchronize()
                best = min(best, a.elapsed_time(b) * 1e-3 / reps)
        row.append("cfg %d: %.1f us (%.1f TF)" % (cfg, best * 1e6, 2 * n ** 3 / best / 1e12))
    print("n=%d  " % n + "   ".join(row), flush=True)
native.gemm_config(0)
