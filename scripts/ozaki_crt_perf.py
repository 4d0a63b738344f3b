"""Modular (CRT) mode of the int8 engine against its digit mode and the fp64 DMMA engine: (1) one 8192^3 fp64-equivalent product
(time with CUDA events, error against the fp64 product), (2) NLL+grad at N = 16384, D = 16 with the products >= 8192 rows on the
engine (time, agreement with the DMMA results; CASES = slices:min_n pairs).  slices <= 8: digits (S (S + 1) / 2 int8 products); >= 10: moduli (one product each).
Writes gpurun_out/ozaki_crt_perf.json."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from bench import synth  # noqa: E402
from gaussian_process_optimization_b200 import native  # noqa: E402

out = {}
n = int(os.environ.get("GEMM_N", "8192"))
g = torch.Generator(device="cuda").manual_seed(1)
A = torch.randn(n, n, generator=g, dtype=torch.float64, device="cuda")
B = torch.randn(n, n, generator=g, dtype=torch.float64, device="cuda")
ref = A @ B.T
C = torch.zeros_like(ref)
gem = {}
for S in [int(a) for a in os.environ.get("GEMM_SLICES", "7,8,16,17,18").split(",")]:
    for _ in range(2):
        native.ozaki_dgemm(0, 0, 1.0, A, B, 0.0, C, slices=S)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        native.ozaki_dgemm(0, 0, 1.0, A, B, 0.0, C, slices=S)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    gem["slices_%d" % S] = {"ms": ms, "tflops_fp64_equivalent": 2.0 * n ** 3 / ms * 1e-9,
                            "max_err_rel_to_max": float((C - ref).abs().max() / ref.abs().max())}
    print("gemm", n, S, gem["slices_%d" % S], flush=True)
out["gemm_%d" % n] = gem
del A, B, C, ref
torch.cuda.empty_cache()

N, D = int(os.environ.get("FIT_N", "16384")), 16
X, Y, ls = synth(N, D)
m = native.NativeModel("rbf", True, D, 1, n_cap=N, cand_block=1024)
m.set_data(torch.from_numpy(X).cuda(), torch.from_numpy(Y).cuda())
fit = {}
base = None
# CASES: slices:min_n pairs (0:0 = the DMMA engine, must come first)
cases = [tuple(int(x) for x in c.split(":")) for c in os.environ.get("CASES", "0:0,7:8192,8:8192,16:8192,17:8192").split(",")]
for S, MIN_N in cases:
    native.set_ozaki(MIN_N, S if S else 8)
    ts = []
    for i in range(4):
        m.set_theta(1.0, ls, 1e-2)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        info, logL, gr = m.fit(True)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    if S == 0:
        base = (logL, gr.copy())
    key = "dmma" if S == 0 else "slices_%d_min_n_%d" % (S, MIN_N)
    fit[key] = {
        "ms": min(ts[1:]) * 1e3, "info": int(info), "logL_rel_vs_dmma": abs(logL - base[0]) / abs(base[0]),
        "grad_rel_vs_dmma": float(np.max(np.abs(gr - base[1])) / np.max(np.abs(base[1])))}
    print("fit", N, key, fit[key], flush=True)
native.set_ozaki(0, 8)
out["fit_%d" % N] = fit
m.close()
os.makedirs("gpurun_out", exist_ok=True)
out["combine_dp4a"] = int(os.environ.get("GPB_OZAKI_COMBINE", "1"))
json.dump(out, open(os.environ.get("OUT", "gpurun_out/ozaki_crt_perf.json"), "w"), indent=1)
