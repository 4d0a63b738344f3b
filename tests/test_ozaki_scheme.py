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
