"""First-look timings on the GPU box (not the bench contract): DMMA GEMM vs cuBLAS DGEMM, NLL+grad at a few sizes."""
import json
import sys
import time

import numpy as np
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
        best = min(best, a.elapsed_time(b) * 1e-3)
    return best


out = {}
for n in (128, 256, 512, 1024):
    A = torch.randn(n, n, dtype=torch.float64, device="cuda")
    B = torch.randn(n, n, dtype=torch.float64, device="cuda")
    C = torch.zeros(n, n, dtype=torch.float64, device="cuda")
    res = {}
    for cfg in (1, 2, 3):
        native.gemm_config(cfg)
        t = ev_time(lambda: native.dgemm(0, 0, 1.0, A, B, 0.0, C), reps=20, warm=3)
        res["cfg%d_us" % cfg] = t * 1e6
    native.gemm_config(0)
    out["gemm_small_%d" % n] = res
    print(n, res, flush=True)
for n in (2048, 4096, 8192):
    A = torch.randn(n, n, dtype=torch.float64, device="cuda")
    B = torch.randn(n, n, dtype=torch.float64, device="cuda")
    C = torch.zeros(n, n, dtype=torch.float64, device="cuda")
    t_cublas = ev_time(lambda: torch.matmul(A, B.T, out=C))
    res = {"cublas_tflops": 2 * n ** 3 / t_cublas / 1e12}
    for ta, tb, name in ((0, 0, "NT"), (0, 1, "NN"), (1, 1, "TN")):
        t = ev_time(lambda: native.dgemm(ta, tb, 1.0, A, B, 0.0, C))
        res["gpb_%s_tflops" % name] = 2 * n ** 3 / t / 1e12
    out["gemm_%d" % n] = res
    print(n, res, flush=True)

for kind, N, D in (("rbf", 4096, 8), ("rbf", 16384, 16), ("mat52", 16384, 16)):
    rs = np.random.RandomState(1234)
    X = rs.uniform(0, 1, (N, D))
    w = rs.randn(D)
    Y = np.sin(X @ w)[:, None] + 0.05 * rs.randn(N, 1)
    Y = (Y - Y.mean()) / Y.std()
    ls = 0.5 + 0.5 * np.arange(D) / D
    m = native.NativeModel(kind, True, D, 1, n_cap=N, cand_block=1024)
    m.set_data(X, Y)
    m.set_theta(1.0, ls, 1e-2)
    ts = []
    for i in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        info, logL, g = m.fit(True)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    flops = N ** 3 + N ** 2 * (6 * D + 62)
    out["fit_%s_%d_%d" % (kind, N, D)] = {"s": min(ts), "tflops": flops / min(ts) / 1e12, "logL": logL, "info": info}
    print(kind, N, D, ts, logL, flops / min(ts) / 1e12, flush=True)
    m.close()
json.dump(out, open("gpurun_out/quick_perf.json", "w"), indent=1)
