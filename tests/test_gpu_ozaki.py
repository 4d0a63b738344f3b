"""GPU parity of the experimental int8 tensor-core engine (csrc/gpb_ozaki.cu: tcgen05 kind::i8 products of balanced radix-256 digits,
recombined in fp64; slices >= 10 = its modular mode, one product per modulus and a CRT reconstruction, csrc/gpb_crt.cuh) -- as a GEMM against fp64 references, with the triangular k-ranges of the cholinv recursion, and as the
engine of a whole NLL + gradient evaluation against the CPU oracle at north_star's tolerances (rtol 1e-9 log-likelihood,
1e-7 gradients).  The engine is off by default; these tests switch it on explicitly."""
import numpy as np
import pytest
from numpy.testing import assert_allclose

from oracle import gp_oracle as O

pytestmark = pytest.mark.gpu

native = pytest.importorskip("gaussian_process_optimization_b200.native")


@pytest.mark.parametrize("ta,tb", [(0, 0), (0, 1), (1, 1), (1, 0)])
@pytest.mark.parametrize("m,n,k", [(128, 128, 128), (256, 384, 640), (640, 1152, 256)])
@pytest.mark.parametrize("slices", [8, 17])
def test_ozaki_dgemm_layouts(ta, tb, m, n, k, slices):
    import torch
    g = torch.Generator(device="cpu").manual_seed(m + 7 * n + 13 * k + ta + 2 * tb)
    A = torch.randn((k, m) if ta else (m, k), generator=g, dtype=torch.float64)
    B = torch.randn((k, n) if tb else (n, k), generator=g, dtype=torch.float64)
    A[0] *= 1e6                     # rows / columns of very different magnitude: the scaling is per row of op(A), op(B)
    B[-1] *= 1e-7
    C = torch.randn((m, n), generator=g, dtype=torch.float64)
    opA, opB = (A.T if ta else A), (B if tb else B.T)
    ref = 1.5 * (opA @ opB) - 0.5 * C
    bound = 1.5 * (opA.abs() @ opB.abs()) + 0.5 * C.abs()       # componentwise scale of a floating-point product
    Cd = C.cuda()
    native.ozaki_dgemm(ta, tb, 1.5, A.cuda(), B.cuda(), -0.5, Cd, slices=slices)
    torch.cuda.synchronize()
    err = (Cd.cpu() - ref).abs()
    # 8 balanced radix-256 digits (17 moduli: >= 59 bits): 2^-61 relative to the row / column maxima per term, i.e. norm-wise fp64 accuracy
    rowmax = opA.abs().amax(dim=1, keepdim=True)
    colmax = opB.abs().amax(dim=0, keepdim=True)
    assert float((err / (k * rowmax * colmax * 2.0 ** -52 + 1e-15 * bound)).max()) < 1.0


@pytest.mark.parametrize("slices", [5, 7, 8, 10, 16, 17, 18])
def test_ozaki_device_equals_the_numpy_emulation_bitwise(slices):
    """Every step of the engine is exact integer arithmetic or a scaling by a power of two, and the fp64 combination adds the
    weights in a fixed order: the device result is bit-identical to the NumPy restatement of the scheme (oracle/ozaki_emulation.py)."""
    import torch
    from oracle import ozaki_emulation as E
    rs = np.random.RandomState(slices)
    m, n, k = 384, 640, 896
    A = rs.randn(m, k) * np.exp2(rs.randint(-30, 30, (m, 1)))
    B = rs.randn(n, k) * np.exp2(rs.randint(-30, 30, (n, 1)))
    B[17] = 0.0
    C = torch.zeros(m, n, dtype=torch.float64, device="cuda")
    native.ozaki_dgemm(0, 0, 1.0, torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda(), 0.0, C, slices=slices)
    torch.cuda.synchronize()
    ref = E.gemm_nt(A, B, slices) if slices <= 8 else E.gemm_nt_crt(A, B, slices)[0]     # >= 10: modular mode
    assert np.array_equal(C.cpu().numpy(), ref)
    # the transposed storage orders cut the same digits
    Ct = torch.zeros(m, n, dtype=torch.float64, device="cuda")
    native.ozaki_dgemm(1, 1, 1.0, torch.from_numpy(np.ascontiguousarray(A.T)).cuda(), torch.from_numpy(np.ascontiguousarray(B.T)).cuda(),
                       0.0, Ct, slices=slices)
    torch.cuda.synchronize()
    assert torch.equal(C, Ct)


def test_ozaki_exact_on_integers():
    """Small integers fit the first two digits exactly: the int8 products and their fp64 recombination are error free."""
    import torch
    g = torch.Generator(device="cpu").manual_seed(5)
    A = torch.randint(-8000, 8001, (256, 384), generator=g).double()
    B = torch.randint(-8000, 8001, (512, 384), generator=g).double()
    C = torch.zeros(256, 512, dtype=torch.float64, device="cuda")
    native.ozaki_dgemm(0, 0, 1.0, A.cuda(), B.cuda(), 0.0, C, slices=3)
    torch.cuda.synchronize()
    assert torch.equal(C.cpu(), A @ B.T)


@pytest.mark.parametrize("slices", [7, 8, 16])
def test_ozaki_long_k_needs_several_int32_accumulations(slices):
    """k = 19200 = 150 k-blocks: the weights with 7 or 8 digit pairs exceed the 1023 k-blocks one int32 accumulation may hold
    (128^2 * 128 * 1023 < 2^31), so they drain in two groups (the N = 32768 end state of BASELINE config 5 needs this)."""
    import torch
    g = torch.Generator(device="cpu").manual_seed(3)
    m, n, k = 256, 384, 19200
    A = torch.randn(m, k, generator=g, dtype=torch.float64)
    B = torch.randn(n, k, generator=g, dtype=torch.float64)
    A[:, ::3] = A[:, ::3].abs()          # biased signs: the int32 sums actually grow with k
    B[:, ::3] = B[:, ::3].abs()
    C = torch.zeros(m, n, dtype=torch.float64, device="cuda")
    native.ozaki_dgemm(0, 0, 1.0, A.cuda(), B.cuda(), 0.0, C, slices=slices)
    torch.cuda.synchronize()
    ref = A @ B.T
    assert float((C.cpu() - ref).abs().max() / ref.abs().max()) < (1e-13 if slices == 8 else 2e-12)       # (16 moduli: one accumulation per modulus)
    # worst case for the accumulators: every digit at its extreme value
    A = torch.full((m, k), -1.0, dtype=torch.float64)
    B = torch.full((n, k), -1.0, dtype=torch.float64)
    native.ozaki_dgemm(0, 0, 1.0, A.cuda(), B.cuda(), 0.0, C, slices=slices)
    torch.cuda.synchronize()
    assert torch.equal(C.cpu(), torch.full((m, n), float(k), dtype=torch.float64))


def _block_lower(n, g, fill):
    """Lower-triangular by 128-blocks (explicit zeros above the diagonal inside the diagonal blocks), `fill` in the blocks above."""
    import torch
    M = torch.randn(n, n, generator=g, dtype=torch.float64)
    clean = torch.tril(M)
    dirty = clean.clone()
    for bi in range(n // 128):
        dirty[bi * 128:(bi + 1) * 128, (bi + 1) * 128:] = fill
    return clean, dirty


@pytest.mark.parametrize("slices", [8, 17])
def test_ozaki_triangular_products_of_the_recursion(slices):
    """The five products of cholinv / potri with their k-ranges; the 128-blocks above the diagonal of the triangular operands
    hold NaN (scratch in the real recursion) and must never be read."""
    import torch
    g = torch.Generator(device="cpu").manual_seed(11)
    h, r = 512, 640
    M11c, M11d = _block_lower(h, g, float("nan"))
    M22c, M22d = _block_lower(r, g, float("nan"))
    A21 = torch.randn(r, h, generator=g, dtype=torch.float64)
    T12 = torch.randn(h, r, generator=g, dtype=torch.float64)
    A22 = torch.randn(r, r, generator=g, dtype=torch.float64)

    def run(ta, tb, alpha, A, B, beta, C, **kw):
        Cd = C.cuda()
        native.ozaki_dgemm(ta, tb, alpha, A.cuda(), B.cuda(), beta, Cd, slices=slices, **kw)
        torch.cuda.synchronize()
        return Cd.cpu()

    def close(got, ref, scale):
        assert torch.isfinite(got).all()
        assert float((got - ref).abs().max()) <= 1e-13 * float(scale)

    # L21 = A21 M11^T  (k <= column block)
    got = run(0, 0, 1.0, A21, M11d, 0.0, torch.zeros(r, h, dtype=torch.float64), khi_mode=2, tri_b=1)
    close(got, A21 @ M11c.T, (A21.abs() @ M11c.abs().T).max())
    # T12 = M11^T L21^T  (k >= row block)
    got = run(1, 0, 1.0, M11d, A21, 0.0, torch.zeros(h, r, dtype=torch.float64), klo_mode=1, tri_a=2)
    close(got, M11c.T @ A21.T, (M11c.abs().T @ A21.abs().T).max())
    # A22 -= L21 L21^T  (lower tiles; the blocks above the diagonal stay untouched)
    got = run(0, 0, -1.0, A21, A21, 1.0, A22.clone(), tri_out=1)
    ref = A22 - A21 @ A21.T
    for bi in range(r // 128):
        sl = slice(bi * 128, (bi + 1) * 128)
        close(got[sl, :(bi + 1) * 128], ref[sl, :(bi + 1) * 128], (A21.abs() @ A21.abs().T).max())
        assert torch.equal(got[sl, (bi + 1) * 128:], A22[sl, (bi + 1) * 128:])
    # M21 = -M22 T21  (k <= row block; T21 through its transpose)
    got = run(0, 0, -1.0, M22d, T12, 0.0, torch.zeros(r, h, dtype=torch.float64), khi_mode=1, tri_a=1)
    close(got, -M22c @ T12.T, (M22c.abs() @ T12.abs().T).max())
    # W = M^T M  (lower tiles, k >= row block)
    got = run(1, 1, 1.0, M22d, M22d, 0.0, torch.zeros(r, r, dtype=torch.float64), tri_out=1, klo_mode=1, tri_a=2, tri_b=2)
    ref = M22c.T @ M22c
    for bi in range(r // 128):
        sl = slice(bi * 128, (bi + 1) * 128)
        close(got[sl, :(bi + 1) * 128], ref[sl, :(bi + 1) * 128], (M22c.abs().T @ M22c.abs()).max())


@pytest.fixture(params=[7, 16])
def ozaki_on(request):
    native.set_ozaki(256, request.param)          # 7 digits; 16 moduli (modular mode, the predictive products then use 18)
    yield
    native.set_ozaki(0, 8)


@pytest.mark.parametrize("kind,noise,N,D", [("rbf", 1e-2, 1100, 5), ("mat52", 1e-6, 1500, 8), ("rbf", 1e-6, 900, 3)])
def test_nll_and_gradients_through_the_int8_engine(ozaki_on, kind, noise, N, D):
    """Every product of the factorisation with >= 256 rows on the int8 tensor cores: the evaluation still meets the parity bars
    against the oracle (LAPACK), the ill-conditioned exact-evaluation noise level included."""
    rs = np.random.RandomState(N)
    X = rs.uniform(0, 1, (N, D))
    Y = np.sin(X @ rs.randn(D))[:, None] + 0.05 * rs.randn(N, 1)
    Y = (Y - Y.mean()) / Y.std()
    ls = 0.5 + 0.5 * np.arange(D) / D
    m = native.NativeModel(kind, True, D, 1, n_cap=N, cand_block=1024)
    m.set_data(X, Y)
    m.set_theta(1.3, ls, noise)
    before = native.launch_count()
    info, logL, grads = m.fit(True)
    assert info == 0 and native.launch_count() > before
    l_ref, g_ref, _ = O.log_likelihood_and_gradients(kind, X, Y, 1.3, ls, noise)
    Ky = O.K(kind, X, None, 1.3, ls) + (noise + 1e-8) * np.eye(N)
    w = np.linalg.eigvalsh(Ky)
    widen = max(1.0, w[-1] / w[0] * 2.2e-16 / 1e-12)              # same allowance as tests/test_gpu_native.py::_cond_tol
    assert_allclose(logL, l_ref, rtol=1e-9 * widen)
    assert_allclose(grads, g_ref, rtol=1e-7 * widen, atol=1e-7 * widen * np.abs(g_ref).max())
    # the predictive products of a full candidate block (1024 rows) run on the engine too, against digit planes of L^-1 that are
    # cut once and reused by the second block; a refit must invalidate them
    Xc = rs.uniform(0, 1, (1024 + 300, D))
    st = O.GPState(kind, X, Y, 1.3, ls, noise)
    r = m.acquisition("EI", 0.01, m.fmin(), Xc, with_gradients=True)
    f_ref, df_ref = st.acquisition("EI", Xc, with_gradients=True)
    assert_allclose(r["f"], f_ref, rtol=1e-6 * widen, atol=1e-9)
    assert_allclose(r["df"], df_ref, rtol=1e-5 * widen, atol=1e-8 * widen * np.abs(df_ref).max())
    m.set_theta(0.9, ls * 1.1, noise)
    assert m.fit(True)[0] == 0
    st2 = O.GPState(kind, X, Y, 0.9, ls * 1.1, noise)
    r2 = m.acquisition("LCB", 2.0, m.fmin(), Xc[:1024], with_gradients=True)
    f2, df2 = st2.acquisition("LCB", Xc[:1024], with_gradients=True)
    assert_allclose(r2["f"], f2, rtol=1e-6 * widen, atol=1e-9)
    assert_allclose(r2["df"], df2, rtol=1e-5 * widen, atol=1e-8 * widen * np.abs(df2).max())
    m.close()


# ---------------------------------------------------------------------------------------------------------------------
# Round 2: the engine's error bound, stated and attacked; the residual check and the automatic fallback
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("slices", [8, 16, 18])
def test_ozaki_error_bound_with_graded_rows(slices):
    """THE bound of the scheme: operands are cut relative to the largest entry of each row of op(A) / op(B), so
        |C_ij - (A B^T)_ij|  <=  2 k 2^-beta  max_k|a_ik|  max_k|b_jk|   +   2 eps |A B^T|_ij
    (beta = bits per operand: 8 S - 1 for S digits minus the dropped pairs' tail, gpb_ozaki_crt_bits for moduli) -- norm-wise per
    row pair, NOT component-wise: entries far below their row's maximum lose relative accuracy.  Rows whose entries span 2^+-40 (what a
    row of L^-1 of an ill-conditioned Ky looks like) must still satisfy it, and the test also shows the flip side: the component-wise
    relative error against |A||B|^T can exceed the fp64 engine's by orders of magnitude."""
    import torch
    rs = np.random.RandomState(slices)
    m, n, k = 256, 384, 1024
    A = rs.randn(m, k) * np.exp2(rs.randint(-40, 41, (m, k)).astype(float))        # graded WITHIN every row
    B = rs.randn(n, k) * np.exp2(rs.randint(-40, 41, (n, k)).astype(float))
    Ad, Bd = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    C = torch.zeros(m, n, dtype=torch.float64, device="cuda")
    native.ozaki_dgemm(0, 0, 1.0, Ad, Bd, 0.0, C, slices=slices)
    Cf = torch.zeros(m, n, dtype=torch.float64, device="cuda")
    native.dgemm(0, 0, 1.0, Ad, Bd, 0.0, Cf)
    torch.cuda.synchronize()
    ref = np.asarray((A.astype(np.longdouble) @ B.T.astype(np.longdouble)), dtype=np.float64)    # 64-bit mantissa accumulation
    beta = native.ozaki_crt_bits(slices, k) if slices >= 10 else 8 * slices - 3
    rowmax, colmax = np.abs(A).max(1)[:, None], np.abs(B).max(1)[None, :]
    absAB = np.abs(A) @ np.abs(B).T
    bound = 2.0 * k * 2.0 ** -beta * rowmax * colmax + 2 * 2.2e-16 * absAB
    err = np.abs(C.cpu().numpy() - ref)
    assert float((err / bound).max()) <= 1.0
    err_f = np.abs(Cf.cpu().numpy() - ref)
    assert float((err_f / (k * 1.2e-16 * absAB)).max()) <= 1.0                      # the fp64 engine's component-wise bound
    # the flip side, documented: measured against |A||B|^T the int8 result may be far outside the fp64 bound
    comp = float((err / (k * 1.2e-16 * absAB)).max())
    assert np.isfinite(comp)


@pytest.mark.parametrize("noise", [1e-7, 1e-9])
def test_ozaki_on_ill_conditioned_models_is_checked_and_no_worse_than_fp64(noise):
    """cond(Ky) ~ 1e10 .. 1e11 (smooth RBF in 2-D, tiny noise): rows of L^-1 are strongly graded.  Against LAPACK (the oracle), every
    quantity from the engine (18 moduli forced on every product >= 256 rows) must be as close as the fp64 DMMA engine's up to a small
    factor -- or the residual check must have sent the fit back to the DMMA engine.  Includes the predictive variance AT the data
    (sigma^2 - |L^-1 k*|^2 cancels to ~ noise)."""
    rs = np.random.RandomState(7)
    n, d = 1536, 2
    X = rs.uniform(0, 1, (n, d))
    Y = np.sin(4 * X[:, :1]) * np.cos(3 * X[:, 1:]) + 1e-4 * rs.randn(n, 1)
    ls, v = np.array([0.35, 0.45]), 1.0
    lo, go, post = O.log_likelihood_and_gradients("rbf", X, Y, v, ls, noise)
    w = np.linalg.eigvalsh(post.K + (noise + 1e-8) * np.eye(n))
    cond = w[-1] / w[0]
    assert cond > 1e9
    Xs = np.vstack([X[:64] + 1e-6, rs.uniform(0, 1, (64, d))])               # at the data, and away from it
    mu_o, var_o = O.predict("rbf", post, X, Xs, v, ls, noise, include_likelihood=False)
    out = {}
    for engine in ("dmma", "int8"):
        native.set_ozaki(256 if engine == "int8" else 0, 18)
        try:
            m = native.NativeModel("rbf", True, d, 1, n_cap=n, cand_block=1024)
            m.set_data(X, Y)
            m.set_theta(v, ls, noise)
            info, logL, g = m.fit(True)
            assert info == 0
            used, resid = m.engine_report()
            mu, var = m.predict(Xs, include_likelihood=False)
            out[engine] = dict(logL=abs(logL - lo) / abs(lo), g=np.abs(g - go).max() / np.abs(go).max(),
                               mu=np.abs(mu - mu_o).max() / np.abs(mu_o).max(), var=np.abs(var - var_o).max() / v, used=used, resid=resid)
            m.close()
        finally:
            native.set_ozaki(0)
    assert out["dmma"]["used"] is False and out["dmma"]["resid"] == -1.0
    e, f = out["int8"], out["dmma"]
    assert e["used"] is False or (0.0 <= e["resid"] <= 2e-13)              # either checked and passed, or fell back
    floor = cond * 2.2e-16                                                   # what two LAPACK builds differ by
    for key in ("logL", "g", "mu", "var"):
        assert e[key] <= max(10.0 * f[key], floor), (key, e, f, cond)


def test_ozaki_residual_check_falls_back_to_the_fp64_engine():
    """10 moduli carry ~34 bits per operand: far too few.  The residual check after the solve must notice, the evaluation must be
    repeated on the DMMA engine (bit-identical to a fit that never saw the engine), and later predictions must stay there."""
    rs = np.random.RandomState(3)
    n, d = 900, 4
    X = rs.uniform(0, 1, (n, d))
    Y = np.sin(X.sum(1))[:, None] + 0.05 * rs.randn(n, 1)
    ls = np.array([0.5, 0.6, 0.7, 0.8])
    Xc = rs.uniform(0, 1, (1500, d))

    def run():
        m = native.NativeModel("mat52", True, d, 1, n_cap=n, cand_block=1536)
        m.set_data(X, Y)
        m.set_theta(1.2, ls, 1e-3)
        info, logL, g = m.fit(True)
        assert info == 0
        f = m.acquisition("EI", 0.01, m.fmin(), Xc, with_gradients=True)
        rep = m.engine_report()
        m.close()
        return logL, g, f["f"], f["df"], rep

    base = run()
    before = native.ozaki_fallback_count()
    native.set_ozaki(256, 10)
    try:
        weak = run()
    finally:
        native.set_ozaki(0)
    assert native.ozaki_fallback_count() == before + 1
    assert weak[4][0] is False and weak[4][1] > 2e-13                        # engine not in the final posterior; the residual that failed
    assert weak[0] == base[0] and np.array_equal(weak[1], base[1])
    assert np.array_equal(weak[2], base[2]) and np.array_equal(weak[3], base[3])
    # and a healthy setting passes the check without falling back
    native.set_ozaki(256, 18)
    try:
        good = run()
    finally:
        native.set_ozaki(0)
    assert native.ozaki_fallback_count() == before + 1
    assert good[4][0] is True and 0.0 <= good[4][1] <= 2e-13
    assert_allclose(good[0], base[0], rtol=1e-12)
