"""Bring-up check of the int8 tensor-core Ozaki engine (csrc/gpb_ozaki.cu) against torch fp64 matmul on the GPU box."""
import json
import os
import sys

import torch

sys.path.insert(0, ".")
from gaussian_process_optimization_b200 import native  # noqa: E402


def ev_time(fn, reps=3, warm=1):
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
g = torch.Generator(device="cuda").manual_seed(0)
# 1. exact small-integer case: one digit must already be exact
m, n, k = 128, 256, 128
A = torch.randint(-100, 101, (m, k), device="cuda", generator=g).double()
B = torch.randint(-100, 101, (n, k), device="cuda", generator=g).double()
C = torch.zeros(m, n, dtype=torch.float64, device="cuda")
native.ozaki_dgemm(0, 0, 1.0, A, B, 0.0, C, slices=2)
torch.cuda.synchronize()
ref = A @ B.t()
err = float((C - ref).abs().max())
print("integer case 128x256x128 S=2: max abs err", err, "ref max", float(ref.abs().max()), flush=True)
out["integer_case_err"] = err
if err != 0.0:
    bad = (C != ref).nonzero()
    print("first mismatches:", bad[:8].tolist(), C[bad[0, 0], bad[0, 1]].item(), ref[bad[0, 0], bad[0, 1]].item(), flush=True)

for (m, n, k) in ((256, 512, 384), (1024, 1024, 1024), (2048, 1920, 4096)):
    for (ta, tb) in ((0, 0), (1, 0), (0, 1), (1, 1)):
        A = torch.randn((k, m) if ta else (m, k), dtype=torch.float64, device="cuda", generator=g)
        B = torch.randn((k, n) if tb else (n, k), dtype=torch.float64, device="cuda", generator=g)
        C0 = torch.randn(m, n, dtype=torch.float64, device="cuda", generator=g)
        opA = A.t() if ta else A
        opB = B if tb else B.t()
        for S in (4, 8):
            for beta in (0.0, 1.0):
                C = C0.clone()
                native.ozaki_dgemm(ta, tb, -0.5, A, B, beta, C, slices=S)
                torch.cuda.synchronize()
                ref = -0.5 * (opA @ opB) + beta * C0
                err = float((C - ref).abs().max() / ref.abs().max())
                key = "m%d_n%d_k%d_ta%d_tb%d_S%d_beta%g" % (m, n, k, ta, tb, S, beta)
                out[key] = err
                print(key, "rel err %.3e" % err, flush=True)

# speed
for n in (4096, 8192):
    A = torch.randn(n, n, dtype=torch.float64, device="cuda", generator=g)
    B = torch.randn(n, n, dtype=torch.float64, device="cuda", generator=g)
    C = torch.zeros(n, n, dtype=torch.float64, device="cuda")
    t_dmma = ev_time(lambda: native.dgemm(0, 0, 1.0, A, B, 0.0, C))
    res = {"dmma_tflops": 2 * n ** 3 / t_dmma / 1e12}
    for S in (6, 7, 8):  # balanced radix-256 digits
        t = ev_time(lambda: native.ozaki_dgemm(0, 0, 1.0, A, B, 0.0, C, slices=S))
        ref = A @ B.t()
        res["ozaki_S%d" % S] = {"ms": t * 1e3, "effective_tflops": 2 * n ** 3 / t / 1e12, "int8_tops": S * (S + 1) / 2 * 2 * n ** 3 / t / 1e12,
                                "rel_err": float((C - ref).abs().max() / ref.abs().max())}
    out["speed_%d" % n] = res
    print(n, res, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/ozaki_engine_check.json", "w"), indent=1)
