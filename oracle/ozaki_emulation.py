"""TEST INFRASTRUCTURE (like everything under oracle/): a NumPy restatement of the digit scheme of the experimental int8
tensor-core engine (gaussian_process_optimization_b200/csrc/gpb_ozaki.cu), exact in fp64 because the digit matrices are small
integers (|digit product sums| < 2^53).  Used by tests/test_ozaki_scheme.py (CPU) and scripts/ozaki_numerics_study.py; the
product path never imports it.  Nothing here follows the reference -- the reference has no such path; the scheme is checked
against plain fp64 / extended-precision products instead."""
import numpy as np


def split_rows_balanced(A, S):
    """A[i, :] = 2^e_i * sum_s D_s[i, :] 2^(1 - 8 s)  (+ the rounding of the last digit, <= 2^(e_i - 8 S)).
    Digits D_s are integers in [-128, 127] (|D_1| <= 65): the row is scaled by a power of two to |x| < 1/2, turned into the
    integer q = rint(x 2^(8 S - 1)) and cut from the low byte up with carries, as oz_split_kernel does with 64-bit integers."""
    assert 1 <= S <= 8
    A = np.asarray(A, dtype=np.float64)
    amax = np.abs(A).max(axis=1, keepdims=True)
    m, e = np.frexp(np.where(amax > 0, amax, 1.0))           # amax = m 2^e, m in [0.5, 1)
    e = np.where(amax > 0, e, 0) + 1                          # |x| 2^-e < 1/2
    q = np.rint(A * np.exp2(8.0 * S - 1.0 - e)).astype(np.int64)
    digits = [None] * S
    for s in range(S, 1, -1):
        d = ((q + 128) & 255) - 128
        digits[s - 1] = d
        q = (q - d) >> 8
    digits[0] = q
    return digits, e.astype(np.float64)


def reconstruct(digits, e):
    S = len(digits)
    acc = np.zeros(digits[0].shape)
    for s in range(S, 0, -1):                                 # smallest terms first
        acc += digits[s - 1] * 2.0 ** (1 - 8 * s)
    return acc * np.exp2(e)


def gemm_nt(A, B, S):
    """A @ B.T through S digits per operand and the digit pairs s + t <= S + 1 (S (S + 1) / 2 exact integer products)."""
    Da, ea = split_rows_balanced(A, S)
    Db, eb = split_rows_balanced(B, S)
    C = np.zeros((A.shape[0], B.shape[0]))
    for w in range(S + 1, 1, -1):                             # smallest terms first, like the kernel's final combination
        acc = np.zeros_like(C)
        for s in range(max(1, w - S), min(S, w - 1) + 1):
            acc += Da[s - 1].astype(np.float64) @ Db[w - s - 1].astype(np.float64).T     # exact: integers below 2^53
        C += acc * 2.0 ** (2 - 8 * w)
    return C * np.exp2(ea) * np.exp2(eb).T
