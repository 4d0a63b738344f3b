"""Round-2 decision data (an EXPERIMENT, not the product path): could fp64 products emulated on the INT8 tensor cores
(Ozaki scheme: error-free slicing into 7-bit digits, int8 x int8 -> int32 products, fp64 recombination) beat the DMMA pipe?

Measures on the GPU box, with LIBRARY kernels only (torch.mm fp64 = cuBLAS DGEMM, torch._int_mm = cuBLASLt IGEMM):
  * the DGEMM rate (the ceiling of everything in libgpb200.so today),
  * the int8 GEMM rate,
  * accuracy and time of a straightforward S-slice emulation of C = A B^T for S = 7..10, on a Gaussian matrix and on a
    triangular inverse factor M = L^-1 of an ill-conditioned covariance (the operand class of this repo's recursion).
Writes gpurun_out/ozaki_probe.json.  Nothing here is imported by the package."""
import json
import os
import sys

import numpy as np
import torch


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


BITS = 7


def split_rows(A, S):
    """Row-scaled 7-bit signed digits: A[i, :] = 2^e_i * sum_s D_s[i, :] 2^(-7 s) (+ remainder below 2^(-7 S))."""
    amax = A.abs().amax(dim=1, keepdim=True).clamp_min(1e-300)
    e = torch.ceil(torch.log2(amax))
    e = torch.where(torch.exp2(e) <= amax, e + 1, e)          # |A| / 2^e < 1 strictly
    R = A * torch.exp2(-e)
    digits = []
    for _ in range(S):
        R = R * float(1 << BITS)
        Dg = torch.trunc(R)
        R = R - Dg
        digits.append(Dg.to(torch.int8))
    return digits, e


def ozaki_gemm_nt(A, B, S):
    """C = A B^T with S digits per operand and the digit pairs s + t <= S + 1 (S (S + 1) / 2 int8 products)."""
    Da, ea = split_rows(A, S)
    Db, eb = split_rows(B, S)
    Dbt = [d.t().contiguous() for d in Db]
    C = torch.zeros(A.shape[0], B.shape[0], dtype=torch.float64, device=A.device)
    for w in range(2, S + 2):                                   # w = s + t (1-based digits)
        acc = None
        for s in range(1, w):
            t = w - s
            if s > S or t > S:
                continue
            P = torch._int_mm(Da[s - 1], Dbt[t - 1])
            acc = P.to(torch.float64) if acc is None else acc + P.to(torch.float64)
        C += acc * float(2.0 ** (-BITS * w))
    return C * torch.exp2(ea) * torch.exp2(eb).t()


def main():
    out = {"device": torch.cuda.get_device_name(0), "bits_per_digit": BITS}
    n = 8192
    g = torch.Generator(device="cuda").manual_seed(1)
    A = torch.randn(n, n, dtype=torch.float64, device="cuda", generator=g)
    B = torch.randn(n, n, dtype=torch.float64, device="cuda", generator=g)
    C = torch.empty(n, n, dtype=torch.float64, device="cuda")
    t = ev_time(lambda: torch.matmul(A, B.t(), out=C), reps=5, warm=2)
    out["dgemm_8192_tflops"] = 2 * n ** 3 / t / 1e12
    print("cuBLAS DGEMM 8192^3: %.1f TFLOP/s" % out["dgemm_8192_tflops"], flush=True)
    for m in (8192, 16384):
        a8 = torch.randint(-127, 128, (m, m), dtype=torch.int8, device="cuda")
        b8 = torch.randint(-127, 128, (m, m), dtype=torch.int8, device="cuda")
        try:
            t8 = ev_time(lambda: torch._int_mm(a8, b8), reps=5, warm=2)
            out["int8_gemm_%d_tops" % m] = 2 * m ** 3 / t8 / 1e12
            print("cuBLASLt int8 GEMM %d^3: %.0f TOP/s" % (m, out["int8_gemm_%d_tops" % m]), flush=True)
        except Exception as ex:  # noqa: BLE001
            out["int8_gemm_%d_error" % m] = repr(ex)
            print("int8 GEMM failed:", ex, flush=True)
        del a8, b8
    if "int8_gemm_8192_tops" not in out:
        return out
    # an ill-conditioned operand of the kind the recursion multiplies: M = L^-1 of a Matern52-like covariance + 1e-6 I
    nn = 4096
    rs = np.random.RandomState(3)
    X = torch.tensor(rs.uniform(0, 1, (nn, 4)), device="cuda")
    r = torch.cdist(X, X) / 0.7
    K = (1 + np.sqrt(5.) * r + 5. / 3. * r * r) * torch.exp(-np.sqrt(5.) * r) + 1e-6 * torch.eye(nn, dtype=torch.float64, device="cuda")
    L = torch.linalg.cholesky(K)
    M = torch.linalg.solve_triangular(L, torch.eye(nn, dtype=torch.float64, device="cuda"), upper=False)
    out["cond_case"] = {"n": nn, "max_abs_M": float(M.abs().max()), "min_diag_L": float(L.diagonal().min())}
    cases = {"gauss_8192": (A, B), "MtM_4096 (Ky^-1 = M^T M, cond ~ 1e8+)": (M.t().contiguous(), M.t().contiguous())}
    for name, (P, Q) in cases.items():
        ref = P @ Q.t()
        # componentwise error scale of a real DGEMM: |P| |Q|^T eps
        scale = (P.abs() @ Q.abs().t())
        res = {}
        for S in (7, 8, 9, 10):
            Cz = ozaki_gemm_nt(P, Q, S)
            err_norm = float((Cz - ref).abs().max() / ref.abs().max())
            err_comp = float(((Cz - ref).abs() / scale.clamp_min(1e-300)).max())
            tt = ev_time(lambda: ozaki_gemm_nt(P, Q, S), reps=2, warm=0)
            res["S%d" % S] = {"int8_gemms": S * (S + 1) // 2, "max_err_over_max_ref": err_norm,
                              "max_err_over_absP_absQ": err_comp, "seconds_unfused_torch": tt,
                              "effective_tflops_unfused": 2 * P.shape[0] * Q.shape[0] * P.shape[1] / tt / 1e12,
                              "effective_tflops_if_only_int8_gemms": out["int8_gemm_8192_tops"] / (S * (S + 1) // 2)}
            print(name, "S=%d" % S, res["S%d" % S], flush=True)
        out[name] = res
        del ref, scale
    return out


if __name__ == "__main__":
    res = main()
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/ozaki_probe.json", "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res))
