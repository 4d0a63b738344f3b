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


# ----------------------------------------------------------------------------------------------------------------------
# Modular splitting ("Ozaki scheme II"; the engine's modular mode, csrc/gpb_crt.cuh: slices in [10, 18]).  One int8 product per
# modulus instead of S (S + 1) / 2 digit-pair products: the operands are scaled to integers A', B' with |A' B'^T| <= P / 2,
# P = product of pairwise coprime moduli <= 256, every modulus gives (A' mod p)(B' mod p)^T with balanced int8 residues and an
# exact int32 accumulation, and the integer product is rebuilt by the Chinese remainder theorem (Garner's mixed-radix form with
# balanced digits, then an evaluation in fp64: three digits at a time exactly as integers, Horner over those groups with one
# rounded product and one rounded sum per step -- the order crt::reconstruct uses).  16 moduli carry what 7 digits (28 products)
# carry; 17-18 moduli what 8 digits (36 products) carry.
# ----------------------------------------------------------------------------------------------------------------------
MODULI = (256, 255, 253, 251, 247, 241, 239, 233, 229, 227, 223, 217, 211, 199, 197, 193, 191, 181, 179, 173)   # pairwise coprime
# (the device engine knows the first 18: crt::modulus in csrc/gpb_crt.cuh)


def _balanced_mod(x, p):
    """x mod p in [-(p // 2), (p - 1) // 2] (fits int8 for p <= 256)."""
    r = np.mod(x, p)
    return np.where(r >= (p + 1) // 2, r - p, r)


def crt_bits(nmod, k):
    """Bits beta per operand: the scaled operands satisfy |Q| <= 2^(beta - 1), so k products stay within 2^(ceil(log2 k) + 2 beta - 2),
    which must not exceed 2^floor(log2 P) / 2 <= P / 2 (exact integer arithmetic; crt::operand_bits does the same)."""
    P = 1
    for p in MODULI[:nmod]:
        P *= int(p)
    fl = P.bit_length() - 1
    lk = (int(k) - 1).bit_length()
    return min((fl + 1 - lk) // 2, 62)


def crt_evaluate(v, moduli):
    """Value of the mixed-radix digits v in fp64: groups of three digits exactly (< 2^24), Horner over the groups from the top."""
    n = len(moduli)
    groups, weights = [], []
    for i in range(0, n, 3):
        g = v[i].astype(np.int64)
        if i + 2 < n:
            g = g + int(moduli[i]) * (v[i + 1].astype(np.int64) + int(moduli[i + 1]) * v[i + 2].astype(np.int64))
        elif i + 1 < n:
            g = g + int(moduli[i]) * v[i + 1].astype(np.int64)
        groups.append(g.astype(np.float64))
        if i + 2 < n:
            weights.append(float(int(moduli[i]) * int(moduli[i + 1]) * int(moduli[i + 2])))
    X = groups[-1]
    for gi in range(len(groups) - 2, -1, -1):
        X = X * weights[gi]                                    # one rounding
        X = X + groups[gi]                                     # one rounding
    return X


def split_rows_integer(A, beta):
    """A[i, :] ~ 2^(e_i - beta) * Q[i, :], Q integer with |Q| < 2^(beta - 1) (row-wise power-of-two scaling, last bit rounded)."""
    A = np.asarray(A, dtype=np.float64)
    amax = np.abs(A).max(axis=1, keepdims=True)
    m, e = np.frexp(np.where(amax > 0, amax, 1.0))
    e = np.where(amax > 0, e, 0) + 1                           # |x| 2^-e < 1/2
    Q = np.rint(A * np.exp2(float(beta) - e)).astype(np.int64)
    return Q, e.astype(np.float64)


def garner_balanced(residues, moduli):
    """Mixed-radix digits v (balanced) with X = v_0 + v_1 p_0 + v_2 p_0 p_1 + ... for the X in (-P/2, P/2) with the given residues."""
    v = []
    for i, p in enumerate(moduli):
        t = residues[i].astype(np.int64)
        for j in range(i):
            inv = pow(int(moduli[j]), -1, int(p))
            t = _balanced_mod((t - v[j]) * inv, p)
        v.append(_balanced_mod(t, p))
    return v


def gemm_nt_crt(A, B, nmod):
    """A @ B.T through nmod int8 products (one per modulus) and a CRT reconstruction; the fp64 Horner evaluation of the
    mixed-radix digits rounds once per step relative to the value itself."""
    k = A.shape[1]
    moduli = MODULI[:nmod]
    beta = min(crt_bits(nmod, k), 62)
    Qa, ea = split_rows_integer(A, beta)
    Qb, eb = split_rows_integer(B, beta)
    residues = []
    for p in moduli:
        Ra = _balanced_mod(Qa, p).astype(np.float64)           # int8 planes on the device
        Rb = _balanced_mod(Qb, p).astype(np.float64)
        S = Ra @ Rb.T                                          # exact: |sum| <= k 128^2 < 2^53 (int32 on the device)
        residues.append(_balanced_mod(S.astype(np.int64), p))
    v = garner_balanced(residues, moduli)
    X = crt_evaluate(v, moduli)
    return X * (np.exp2(ea - beta) * np.exp2(eb - beta).T), beta
