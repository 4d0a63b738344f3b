"""GEMM tile-configuration sweep against cuBLAS DGEMM (not the bench contract)."""
import sys

import torch

sys.path.insert(0, ".")
from gaussian_process_optimization_b200 import native  # noqa: E402


def ev_time(fn, reps=4, warm=2):
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
        best = min(best, a.elapsed_time(b) * 1e-3)
    return best


cfgs = [int(c) for c in sys.argv[1].split(",")] if len(sys.argv) > 1 else [1, 4, 5]
for n in (4096, 8192):
    A = torch.randn(n, n, dtype=torch.float64, device="cuda")
    B = torch.randn(n, n, dtype=torch.float64, device="cuda")
    C = torch.zeros(n, n, dtype=torch.float64, device="cuda")
    t = ev_time(lambda: torch.matmul(A, B.T, out=C))
    print("n=%d cuBLAS %.2f TF" % (n, 2 * n ** 3 / t / 1e12), flush=True)
    ref = None
    for cfg in cfgs:
        native.gemm_config(cfg)
        row = []
        for ta, tb, name in ((0, 0, "NT"), (0, 1, "NN"), (1, 1, "TN")):
            t = ev_time(lambda: native.dgemm(ta, tb, 1.0, A, B, 0.0, C))
            row.append("%s %.2f" % (name, 2 * n ** 3 / t / 1e12))
        native.dgemm(0, 0, 1.0, A, B, 0.0, C)
        torch.cuda.synchronize()
        if ref is None:
            ref = C.clone()
        print("  cfg %d: %s TF   max|diff vs first cfg| %.3e" % (cfg, "  ".join(row), float((C - ref).abs().max())), flush=True)
    native.gemm_config(0)
