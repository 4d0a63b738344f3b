"""CPU checks of the digit scheme behind the experimental int8 engine (csrc/gpb_ozaki.cu), on its NumPy restatement
(oracle/ozaki_emulation.py): digit ranges, exact reconstruction, the error of an emulated product against an extended-precision
one, exactness on integers.  The device side is covered by tests/test_gpu_ozaki.py."""
import numpy as np
import pytest

from oracle import ozaki_emulation as E


@pytest.mark.parametrize("S", [1, 3, 7, 8])
def test_digits_are_int8_and_reconstruct_the_operand(S):
    rs = np.random.RandomState(S)
    A = rs.randn(37, 53) * np.exp2(rs.randint(-40, 40, (37, 1)))          # rows of very different magnitude
    A[3] = 0.0                                                            # an all-zero row
    A[5, 7] = -A[5].__abs__().max() * 2                                   # the row maximum is negative
    A[9] = np.exp2(-3)                                                    # exact powers of two (frexp boundary)
    digits, e = E.split_rows_balanced(A, S)
    assert len(digits) == S
    for s, d in enumerate(digits):
        assert d.dtype == np.int64 and d.min() >= -128 and d.max() <= 127
        if s == 0:
            assert np.abs(d).max() <= 65
    back = E.reconstruct(digits, e)
    # only the last digit rounds: half a unit of 2^(1 - 8 S) relative to 2^e
    assert np.all(np.abs(back - A) <= 0.5 * np.exp2(e + 1 - 8 * S) * (1 + 1e-12))
    if S == 8:
        assert np.array_equal(back[np.abs(A) >= np.exp2(e - 9)], A[np.abs(A) >= np.exp2(e - 9)])   # 61 bits: exact for entries near the maximum


def test_emulated_product_accuracy():
    rs = np.random.RandomState(1)
    m, n, k = 24, 31, 600
    A, B = rs.randn(m, k), rs.randn(n, k)
    ref = (A.astype(np.longdouble) @ B.astype(np.longdouble).T).astype(np.float64)
    scale = np.abs(A).max(1, keepdims=True) * np.abs(B).max(1, keepdims=True).T * k
    plain = np.abs(A @ B.T - ref).max()
    for S, bits in ((6, 45), (7, 53), (8, 61)):
        err = np.abs(E.gemm_nt(A, B, S) - ref)
        assert np.all(err <= 8 * scale * 2.0 ** -bits + 4 * np.finfo(float).eps * np.abs(ref)), S
    assert np.abs(E.gemm_nt(A, B, 8) - ref).max() <= 4 * plain + 1e-300     # 8 digits: as good as the plain fp64 product


def test_emulated_product_is_exact_on_integers():
    rs = np.random.RandomState(2)
    A = rs.randint(-8000, 8001, (16, 300)).astype(np.float64)
    B = rs.randint(-8000, 8001, (20, 300)).astype(np.float64)
    assert np.array_equal(E.gemm_nt(A, B, 3), A @ B.T)


def test_int32_accumulation_bound():
    """The kernel chains at most 1023 k-blocks of 128 into one int32 accumulation: the extreme digit product 128 * 128 fits."""
    assert 128 * 128 * 128 * 1023 < 2 ** 31
    assert 128 * 128 * 128 * 1024 >= 2 ** 31


# ---- modular splitting (studied for the next round; CPU emulation only) ----------------------------------------------------------
def test_moduli_are_pairwise_coprime_and_fit_int8():
    from math import gcd
    M = E.MODULI
    assert all(p <= 256 for p in M)
    assert all(gcd(M[i], M[j]) == 1 for i in range(len(M)) for j in range(i))
    x = np.arange(-70000, 70000)
    for p in M:
        r = E._balanced_mod(x, p)
        assert r.min() >= -128 and r.max() <= 127 and np.all((x - r) % p == 0)


def test_garner_rebuilds_signed_integers_exactly():
    rs = np.random.RandomState(0)
    moduli = E.MODULI[:6]
    P = int(np.prod([int(p) for p in moduli], dtype=object))
    X = np.array([rs.randint(-(P // 2) + 1, P // 2) for _ in range(200)] + [0, 1, -1, P // 2 - 1, -(P // 2) + 1], dtype=object)
    residues = [np.array([int(E._balanced_mod(np.int64(int(x) % p), p)) for x in X], dtype=np.int64) for p in moduli]
    v = E.garner_balanced(residues, moduli)
    rebuilt, radix = np.zeros(len(X), dtype=object), 1
    for vi, p in zip(v, moduli):
        assert vi.min() >= -(p // 2) and vi.max() <= (p - 1) // 2
        rebuilt = rebuilt + vi.astype(object) * radix
        radix *= int(p)
    assert np.all(rebuilt == X)


@pytest.mark.parametrize("nmod,bits", [(12, 38), (16, 54), (18, 61)])
def test_crt_product_accuracy(nmod, bits):
    rs = np.random.RandomState(nmod)
    m, n, k = 20, 27, 700
    A = rs.randn(m, k) * np.exp2(rs.randint(-20, 20, (m, 1)))
    B = rs.randn(n, k)
    ref = (A.astype(np.longdouble) @ B.astype(np.longdouble).T).astype(np.float64)
    C, beta = E.gemm_nt_crt(A, B, nmod)
    assert beta >= bits
    scale = np.abs(A).max(1, keepdims=True) * np.abs(B).max(1, keepdims=True).T * k
    assert np.all(np.abs(C - ref) <= 8 * scale * 2.0 ** -(beta - 1) + 4 * np.finfo(float).eps * np.abs(ref))
    if nmod == 16:
        # 16 moduli (16 int8 products) are at least as accurate as 7 balanced digits (28 products)
        assert np.abs(C - ref).max() <= 2 * np.abs(E.gemm_nt(A, B, 7) - ref).max() + 1e-300
