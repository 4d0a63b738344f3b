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


# ---- modular splitting (the engine's modular mode): the restatement itself --------------------------------------------------------
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


# ---- the engine's modular mode: its integer arithmetic (csrc/gpb_crt.cuh) compiled for the host, against the restatement above ---
def _lib():
    from gaussian_process_optimization_b200 import _lib as L
    return L.load()


def test_crt_operand_bits_agree_with_the_library():
    lib = _lib()
    for nmod in range(10, 19):
        for k in (128, 129, 700, 1024, 4096, 16384, 32768, 130944):
            assert lib.gpb_ozaki_crt_bits(nmod, k) == E.crt_bits(nmod, k), (nmod, k)
    assert E.crt_bits(16, 16384) == 56 and E.crt_bits(17, 16384) == 59 and E.crt_bits(18, 16384) == 62


def _host_residues(A, nmod, beta):
    lib = _lib()
    A = np.ascontiguousarray(A, dtype=np.float64)
    planes = np.zeros((nmod,) + A.shape, dtype=np.int8)
    scale = np.zeros(A.shape[0])
    assert lib.gpb_ozaki_crt_host_residues(A.ctypes.data, A.shape[0], A.shape[1], nmod, beta, planes.ctypes.data, scale.ctypes.data) == 0
    return planes, scale


@pytest.mark.parametrize("nmod", [10, 16, 17, 18])
def test_crt_library_residues_equal_the_restatement(nmod):
    """dp4a-style residue extraction (two's-complement bytes weighted by 2^(8k) mod p, one multiply-high reduction) == Q mod p."""
    rs = np.random.RandomState(nmod)
    A = rs.randn(37, 300) * np.exp2(rs.randint(-30, 30, (37, 1)))
    A[3] = 0.0
    A[5, :7] = [1.0, -1.0, 0.5, -0.5, 2.0 ** -40, -2.0 ** -40, 0.0]
    beta = E.crt_bits(nmod, 300)
    planes, scale = _host_residues(A, nmod, beta)
    Q, e = E.split_rows_integer(A, beta)
    assert np.array_equal(scale, np.exp2(e[:, 0] - beta))
    for i, p in enumerate(E.MODULI[:nmod]):
        assert np.array_equal(planes[i].astype(np.int64), E._balanced_mod(Q, p)), p
    # extreme magnitudes of the scaled integer, both signs
    ext = np.array([[0.5 - 2.0 ** -54, -(0.5 - 2.0 ** -54), 2.0 ** -62, -2.0 ** -62, 0.25, -0.25, 0.4999, -0.4999]])
    planes, _ = _host_residues(ext, nmod, 62)
    Q, _ = E.split_rows_integer(ext, 62)
    for i, p in enumerate(E.MODULI[:nmod]):
        assert np.array_equal(planes[i].astype(np.int64), E._balanced_mod(Q, p)), p


@pytest.mark.parametrize("nmod", [10, 11, 12, 13, 14, 15, 16, 17, 18])
def test_crt_library_reconstruction_is_bit_identical_to_the_restatement(nmod):
    """int32 sums -> balanced residues -> Garner digits (folded compile-time constants) -> grouped Horner: the library's host build of
    the device code gives the same doubles as gemm_nt_crt, bit for bit."""
    lib = _lib()
    rs = np.random.RandomState(100 + nmod)
    m, n, k = 24, 40, 1500
    A = rs.randn(m, k) * np.exp2(rs.randint(-10, 10, (m, 1)))
    B = rs.randn(n, k) * np.exp2(rs.randint(-10, 10, (n, 1)))
    beta = E.crt_bits(nmod, k)
    Ra, sa = _host_residues(A, nmod, beta)
    Rb, sb = _host_residues(B, nmod, beta)
    sums = np.stack([Ra[i].astype(np.int64) @ Rb[i].astype(np.int64).T for i in range(nmod)])
    assert np.abs(sums).max() < 2 ** 31
    sums = np.ascontiguousarray(sums.astype(np.int32))
    X = np.zeros((m, n))
    assert lib.gpb_ozaki_crt_host_combine(sums.ctypes.data, m * n, nmod, X.ctypes.data) == 0
    C = X * (sa[:, None] * sb[None, :])
    ref, beta2 = E.gemm_nt_crt(A, B, nmod)
    assert beta2 == beta
    assert np.array_equal(C, ref)
    exact = (A.astype(np.longdouble) @ B.astype(np.longdouble).T).astype(np.float64)
    scale = np.abs(A).max(1, keepdims=True) * np.abs(B).max(1, keepdims=True).T * k
    assert np.all(np.abs(C - exact) <= 8 * scale * 2.0 ** -(beta - 1) + 4 * np.finfo(float).eps * np.abs(exact))


def test_crt_library_reconstruction_at_the_int32_extremes():
    """Residue of an int32 accumulation near +-2^31 (hi * (2^16 mod p) + lo, one reduction)."""
    lib = _lib()
    nmod = 16
    s = np.array([2 ** 31 - 1, -2 ** 31 + 1, 2 ** 31 - 65536, -2 ** 31 + 65536, 65535, -65536, 0, 1, -1, 123456789, -987654321],
                 dtype=np.int64)
    # choose the same X for every modulus: sums_i = s (a consistent residue system of the integer s itself)
    sums = np.ascontiguousarray(np.tile(s.astype(np.int32), (nmod, 1)))
    X = np.zeros(len(s))
    assert lib.gpb_ozaki_crt_host_combine(sums.ctypes.data, len(s), nmod, X.ctypes.data) == 0
    assert np.array_equal(X, s.astype(np.float64))


def test_set_ozaki_accepts_digits_or_moduli_only():
    """slices 1..8 = digits, 10..18 = moduli (modular mode); anything else is refused at the C ABI (no GPU needed)."""
    lib = _lib()
    try:
        for s in (1, 7, 8, 10, 16, 18):
            assert lib.gpb_set_ozaki(8192, s) == 0
        for s in (0, 9, 19, -3):
            assert lib.gpb_set_ozaki(8192, s) != 0
        assert lib.gpb_set_ozaki(-1, 8) != 0
    finally:
        assert lib.gpb_set_ozaki(0, 8) == 0
