"""Round-2 decision data (CPU EXPERIMENT, not the product path): would the parity bars of north_star (rtol 1e-9 on the
log-likelihood, 1e-7 on gradients) survive if every GEMM of the cholinv recursion ran as an INT8
Ozaki emulation with S 7-bit digits per operand instead of fp64 DMMA?

The emulation is simulated EXACTLY in fp64 on the CPU: the digit matrices are small integers, so their products are exact in
fp64 (k * 127^2 < 2^53), which is what int8 x int8 -> int32 tensor-core products would return.  The recursion, leaves
(LAPACK on 128 x 128 blocks) and the downstream formulas are this repo's; the comparison target is the oracle (LAPACK path).

    python scripts/ozaki_numerics_study.py [N]        # writes profiles/r1k_ozaki_numerics_study.json (r1i: before the engine's
                                                      # modular mode existed; r1k: with its operand bits and evaluation order)
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import gp_oracle as O  # noqa: E402
from oracle import ozaki_emulation as E  # noqa: E402

BITS = 7
TILE = 128


def split_rows(A, S):
    amax = np.maximum(np.abs(A).max(axis=1, keepdims=True), 1e-300)
    e = np.ceil(np.log2(amax))
    e = np.where(np.exp2(e) <= amax, e + 1, e)
    R = A * np.exp2(-e)
    digits = []
    for _ in range(S):
        R = R * float(1 << BITS)
        Dg = np.trunc(R)
        R = R - Dg
        digits.append(Dg)
    return digits, e


def make_gemm_balanced(S):
    """The scheme of csrc/gpb_ozaki.cu: balanced radix-256 digits (oracle/ozaki_emulation.py)."""
    return lambda A, B: E.gemm_nt(A, B, S)


def make_gemm(S):
    if S >= 100:                      # modular splitting with S - 100 moduli (the engine's modular mode, csrc/gpb_crt.cuh)
        return lambda A, B: E.gemm_nt_crt(A, B, S - 100)[0]
    if S < 0:
        return make_gemm_balanced(-S)
    """gemm_nt(A, B) = A @ B.T, exact fp64 (S = 0) or the S-digit emulation with the digit pairs s + t <= S + 1."""
    if S == 0:
        return lambda A, B: A @ B.T
    def gemm_nt(A, B):
        Da, ea = split_rows(A, S)
        Db, eb = split_rows(B, S)
        C = np.zeros((A.shape[0], B.shape[0]))
        for w in range(S + 1, 1, -1):                     # smallest terms first
            acc = np.zeros_like(C)
            for s in range(1, w):
                t = w - s
                if s <= S and t <= S:
                    acc += Da[s - 1] @ Db[t - 1].T        # exact: integer-valued, < 2^53
            C += acc * 2.0 ** (-BITS * w)
        return C * np.exp2(ea) * np.exp2(eb).T
    return gemm_nt


def cholinv(A, gemm_nt):
    """L, M = L^-1 by the 2 x 2 recursion of gpb_chol.cu (every product above the 128-leaves through gemm_nt)."""
    n = A.shape[0]
    if n <= TILE:
        L = np.linalg.cholesky(A)
        return L, np.linalg.solve(L, np.eye(n))
    h = ((n // TILE) // 2) * TILE
    L11, M11 = cholinv(A[:h, :h], gemm_nt)
    L21 = gemm_nt(A[h:, :h], M11)                          # A21 M11^T
    A22 = A[h:, h:] - gemm_nt(L21, L21)
    L22, M22 = cholinv(A22, gemm_nt)
    T21 = gemm_nt(L21, M11.T.copy())                       # L21 M11
    M21 = -gemm_nt(M22, T21.T.copy())                      # -M22 T21
    L = np.zeros_like(A)
    M = np.zeros_like(A)
    L[:h, :h], L[h:, :h], L[h:, h:] = L11, L21, L22
    M[:h, :h], M[h:, :h], M[h:, h:] = M11, M21, M22
    return L, M


def evaluate(kind, X, Y, var, ls, noise, S):
    gemm_nt = make_gemm(S)
    Kmat = O.K(kind, X, None, var, ls)
    Ky = Kmat + (noise + 1e-8) * np.eye(X.shape[0])
    L, M = cholinv(Ky, gemm_nt)
    W = gemm_nt(M.T.copy(), M.T.copy())                    # Ky^-1 = M^T M
    alpha = M.T @ (M @ Y)
    logdet = 2 * np.log(np.diag(L)).sum()
    n = X.shape[0]
    logL = 0.5 * (-n * np.log(2 * np.pi) - logdet - float((alpha * Y).sum()))
    dL_dK = 0.5 * (alpha @ alpha.T - W)
    gk = O.update_gradients_full(kind, dL_dK, X, None, var, ls)
    g = np.concatenate([np.atleast_1d(gk[0]), np.atleast_1d(gk[1]).ravel(), [np.trace(dL_dK)]])
    return logL, g


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    D = 8
    rs = np.random.RandomState(1234)
    X = rs.uniform(0, 1, (N, D))
    w = rs.randn(D)
    Y = np.sin(X @ w)[:, None] + 0.05 * rs.randn(N, 1)
    Y = (Y - Y.mean()) / Y.std()
    ls = 0.5 + 0.5 * np.arange(D) / D
    out = {"N": N, "D": D, "bits_per_digit": BITS, "note": "relative deviations from the oracle (LAPACK); S = 0 is the recursion in plain fp64"}
    for kind, noise in (("rbf", 1e-2), ("mat52", 1e-2), ("mat52", 1e-6), ("rbf", 1e-6)):
        l_ref, g_ref, _ = O.log_likelihood_and_gradients(kind, X, Y, 1.0, ls, noise)
        Ky = O.K(kind, X, None, 1.0, ls) + (noise + 1e-8) * np.eye(N)
        ev = np.linalg.eigvalsh(Ky)
        case = {"cond_Ky": float(ev[-1] / ev[0])}
        for S in (0, 7, 8, 9, -6, -7, -8, 115, 116, 117, 118):
            try:
                l, g = evaluate(kind, X, Y, 1.0, ls, noise, S)
                case["S%d" % S] = {"scheme": ("plain fp64" if S == 0 else "modular splitting (CRT), %d moduli" % (S - 100) if S >= 100
                                              else "7-bit truncated digits" if S > 0 else "balanced radix-256 digits"),
                                   "int8_gemms_per_product": (S - 100) if S >= 100 else abs(S) * (abs(S) + 1) // 2,
                                   "logL_rel": float(abs(l - l_ref) / abs(l_ref)),
                                   "grad_rel_max": float(np.max(np.abs(g - g_ref) / np.maximum(np.abs(g_ref), 1e-300))),
                                   "grad_rel_to_norm": float(np.max(np.abs(g - g_ref)) / np.max(np.abs(g_ref)))}
            except np.linalg.LinAlgError as ex:
                case["S%d" % S] = {"error": str(ex)}
            print(kind, noise, "S=%d" % S, case["S%d" % S], flush=True)
        out["%s_noise%g" % (kind, noise)] = case
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles",
                        os.environ.get("STUDY_OUT", "r1k_ozaki_numerics_study.json"))
    with open(path, "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
