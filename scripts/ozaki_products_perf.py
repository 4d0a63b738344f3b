"""Per-product timing of the int8 engine on the five product shapes of the top recursion level (h = r = N / 2) and Ky^-1 = M^T M."""
import json
import os
import sys

import torch

sys.path.insert(0, ".")
from gaussian_process_optimization_b200 import native  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
S = int(sys.argv[2]) if len(sys.argv) > 2 else 7
h = N // 2


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
        best = min(best, a.elapsed_time(b))
    return best


g = torch.Generator(device="cuda").manual_seed(0)
M11 = torch.tril(torch.randn(h, h, dtype=torch.float64, device="cuda", generator=g))
A21 = torch.randn(h, h, dtype=torch.float64, device="cuda", generator=g)
C = torch.zeros(h, h, dtype=torch.float64, device="cuda")
out = {"N": N, "S": S, "dbg": os.environ.get("GPB_OZAKI_DBG", "0")}
out["full_%d" % h] = ev_time(lambda: native.ozaki_dgemm(0, 0, 1.0, A21, A21, 0.0, C, slices=S))
out["L21 (khi=col)"] = ev_time(lambda: native.ozaki_dgemm(0, 0, 1.0, A21, M11, 0.0, C, slices=S, khi_mode=2, tri_b=1))
out["T12 (klo=row)"] = ev_time(lambda: native.ozaki_dgemm(1, 0, 1.0, M11, A21, 0.0, C, slices=S, klo_mode=1, tri_a=2))
out["A22 (tri_out, beta=1)"] = ev_time(lambda: native.ozaki_dgemm(0, 0, -1.0, A21, A21, 1.0, C, slices=S, tri_out=1))
out["M21 (khi=row)"] = ev_time(lambda: native.ozaki_dgemm(0, 0, -1.0, M11, A21, 0.0, C, slices=S, khi_mode=1, tri_a=1))
del A21, C
M = torch.tril(torch.randn(N, N, dtype=torch.float64, device="cuda", generator=g))
W = torch.zeros(N, N, dtype=torch.float64, device="cuda")
out["MtM %d (tri_out, klo=row)" % N] = ev_time(lambda: native.ozaki_dgemm(1, 1, 1.0, M, M, 0.0, W, slices=S, tri_out=1, klo_mode=1, tri_a=2, tri_b=2))
for k, v in out.items():
    print(k, v if isinstance(v, (str, int)) else "%.3f ms" % v, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/ozaki_products_perf_dbg%s.json" % out["dbg"], "w"), indent=1)
