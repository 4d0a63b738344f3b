"""GPU parity tests through the C ABI (libgpb200.so via gaussian_process_optimization_b200.native).

Every check compares the CUDA path with (i) the golden vectors produced by the reference's own source files
(tests/golden/*.npz) and (ii) the CPU oracle (oracle/gp_oracle.py) on seeded inputs.
Tolerances follow BASELINE.json north_star: rtol 1e-9 on K, log-likelihood and predictions, 1e-7 on gradients; quantities
whose conditioning is cond(Ky) * eps (alpha, Ky^-1 and what is built from them, the noise-free variance near the data) carry
that factor explicitly -- two LAPACK builds do not agree better than that either.
"""
import numpy as np
import pytest
from numpy.testing import assert_allclose

from oracle import gp_oracle as O

pytestmark = pytest.mark.gpu

native = pytest.importorskip("gaussian_process_optimization_b200.native")


def _theta(g):
    return g["variance"], g["lengthscale"], g["noise"]


def _cond_tol(g, base):
    """base * max(1, cond(Ky) * eps / 1e-12): widen only for the nearly singular exact-evaluation case."""
    K = g["K"]
    w = np.linalg.eigvalsh(K + (g["noise"] + 1e-8) * np.eye(K.shape[0]))
    cond = w[-1] / w[0]
    return base * max(1.0, cond * 2.2e-16 / 1e-12)


# ---------------------------------------------------------------------------------------------------------------------
# DMMA GEMM engine
# ---------------------------------------------------------------------------------------------------------------------
@pytest.fixture
def gemm_cfg(request):
    """Force one GEMM tile configuration (1 = 64x128, 2 = 64x64, 3 = 32x32 CTA tiles; 0 = automatic) for a test."""
    native.gemm_config(request.param)
    yield request.param
    native.gemm_config(0)


@pytest.mark.parametrize("gemm_cfg", [0, 1, 2, 3], indirect=True)
@pytest.mark.parametrize("ta,tb", [(0, 0), (0, 1), (1, 1), (1, 0)])
@pytest.mark.parametrize("m,n,k", [(128, 128, 16), (256, 384, 272), (512, 128, 1024)])
def test_dgemm_layouts(ta, tb, m, n, k, gemm_cfg):
    import torch
    g = torch.Generator(device="cpu").manual_seed(m + 7 * n + 13 * k + ta + 2 * tb)
    A = torch.randn((k, m) if ta else (m, k), generator=g, dtype=torch.float64)
    B = torch.randn((k, n) if tb else (n, k), generator=g, dtype=torch.float64)
    C = torch.randn((m, n), generator=g, dtype=torch.float64)
    ref = 1.5 * ((A.T if ta else A) @ (B if tb else B.T)) - 0.5 * C
    Cd = C.cuda()
    native.dgemm(ta, tb, 1.5, A.cuda(), B.cuda(), -0.5, Cd)
    torch.cuda.synchronize()
    assert_allclose(Cd.cpu().numpy(), ref.numpy(), rtol=1e-12, atol=1e-11)


@pytest.mark.parametrize("gemm_cfg", [1, 2, 3], indirect=True)
def test_pdinv_each_gemm_config(gemm_cfg):
    """The triangular k-ranges and the lower-tile enumeration of every tile configuration (n = 5 blocks of 128)."""
    n = 600
    rs = np.random.RandomState(n)
    B = rs.randn(n, n + 3)
    A = B @ B.T + 0.5 * n * np.eye(n)
    rc, Ai, L, Li, logdet = native.pdinv(A)
    assert rc == 0
    Lr = np.linalg.cholesky(A)
    assert_allclose(L, Lr, rtol=1e-10, atol=1e-12 * np.abs(Lr).max())
    assert_allclose(Li, np.linalg.inv(Lr), rtol=1e-9, atol=1e-12)
    assert_allclose(Ai, np.linalg.inv(A), rtol=1e-9, atol=1e-12 * np.abs(Ai).max())


def test_dgemm_beta_zero_ignores_nan_in_c():
    import torch
    A = torch.randn((128, 32), dtype=torch.float64)
    B = torch.randn((128, 32), dtype=torch.float64)
    Cd = torch.full((128, 128), float("nan"), dtype=torch.float64, device="cuda")
    native.dgemm(0, 0, 1.0, A.cuda(), B.cuda(), 0.0, Cd)
    assert_allclose(Cd.cpu().numpy(), (A @ B.T).numpy(), rtol=1e-12, atol=1e-12)


# ---------------------------------------------------------------------------------------------------------------------
# (a) kernels
# ---------------------------------------------------------------------------------------------------------------------
def test_kernel_matrix_golden(golden):
    g = golden
    v, ls, _ = _theta(g)
    Ks = native.kern_K(g["kind"], g["X"], None, v, ls)
    assert_allclose(Ks, g["K"], rtol=1e-9, atol=1e-300)
    assert np.array_equal(np.diag(Ks), np.full(Ks.shape[0], v))          # r_ii forced to 0 (stationary.py:164)
    assert_allclose(native.kern_K(g["kind"], g["Xs"], g["X"], v, ls), g["K_cross"], rtol=1e-9)


def test_update_gradients_full_golden(golden):
    g = golden
    v, ls, _ = _theta(g)
    dv, dl = native.kern_update_gradients_full(g["kind"], g["G_sq"], g["X"], None, v, ls)
    assert_allclose(dv, g["ugf_sq_var"].ravel()[0], rtol=1e-7)
    assert_allclose(dl, g["ugf_sq_len"], rtol=1e-7)
    dv, dl = native.kern_update_gradients_full(g["kind"], g["G_rect"], g["Xs"], g["X"], v, ls)
    assert_allclose(dv, g["ugf_rect_var"].ravel()[0], rtol=1e-7)
    assert_allclose(dl, g["ugf_rect_len"], rtol=1e-7)


def test_gradients_X_golden(golden):
    g = golden
    v, ls, _ = _theta(g)
    if not g["ard"]:
        ls = np.full(g["X"].shape[1], ls[0])    # gradients_X divides by lengthscale**2 per column either way
    scale = np.abs(g["gX_sq"]).max()
    assert_allclose(native.kern_gradients_X(g["kind"], g["G_sq"], g["X"], None, v, ls), g["gX_sq"], rtol=1e-7, atol=1e-9 * scale)
    scale = np.abs(g["gX_rect"]).max()
    assert_allclose(native.kern_gradients_X(g["kind"], g["G_rect"], g["Xs"], g["X"], v, ls), g["gX_rect"], rtol=1e-7,
                    atol=1e-9 * scale)


@pytest.mark.parametrize("kind", ["rbf", "mat52"])
def test_kernel_shapes_like_reference_cython_tests(kind):
    """GPy/GPy/testing/cython_tests.py:39-67 shapes: X 300x10, Z 20x10, random dL_dK, square and rectangular."""
    rs = np.random.RandomState(5)
    X, Z = rs.randn(300, 10), rs.randn(20, 10)
    ls = 0.8 + rs.rand(10)
    G1, G2 = rs.randn(300, 300), rs.randn(300, 20)
    assert_allclose(native.kern_K(kind, X, Z, 1.7, ls), O.K(kind, X, Z, 1.7, ls), rtol=1e-9, atol=1e-300)
    for G, X2 in ((G1, None), (G2, Z)):
        dv, dl = native.kern_update_gradients_full(kind, G, X, X2, 1.7, ls)
        rv, rl = O.update_gradients_full(kind, G, X, X2, 1.7, ls)
        assert_allclose(dv, rv, rtol=1e-7)
        assert_allclose(dl, rl, rtol=1e-7)
        ref = O.gradients_X(kind, G, X, X2, 1.7, ls)
        assert_allclose(native.kern_gradients_X(kind, G, X, X2, 1.7, ls), ref, rtol=1e-7, atol=1e-9 * np.abs(ref).max())


@pytest.mark.parametrize("kind", ["rbf", "mat52"])
@pytest.mark.parametrize("rows,n,d", [(64, 128, 16), (71, 389, 5), (203, 1031, 16), (130, 517, 27), (96, 130, 1)])
def test_gradients_X_tiled_kernel_ragged_shapes(kind, rows, n, d):
    """More than 8 output rows take the shared-memory tiled cluster kernel (gpb_kernels.cu gradx_tile_kernel): rows not a multiple of 8,
    points not a multiple of the 128-point chunk or of the 4 slices, every register cap; rectangular and symmetric (tmp + tmp.T)
    forms of stationary.py:354-366 against the oracle, and row-wise against the one-warp-group-per-row kernel (<= 8 rows)."""
    rs = np.random.RandomState(rows + n + d)
    X, Z = rs.rand(rows, d), rs.rand(n, d)
    ls = 0.3 + rs.rand(d)
    G = rs.randn(rows, n)
    ref = O.gradients_X(kind, G, X, Z, 1.3, ls)
    got = native.kern_gradients_X(kind, G, X, Z, 1.3, ls)
    assert_allclose(got, ref, rtol=1e-9, atol=1e-12 * np.abs(ref).max())
    small = np.vstack([native.kern_gradients_X(kind, G[a:a + 8], X[a:a + 8], Z, 1.3, ls) for a in range(0, rows, 8)])
    assert_allclose(got, small, rtol=1e-11, atol=1e-13 * np.abs(ref).max())
    Gs = rs.randn(rows, rows)
    ref = O.gradients_X(kind, Gs, X, None, 1.3, ls)
    assert_allclose(native.kern_gradients_X(kind, Gs, X, None, 1.3, ls), ref, rtol=1e-9, atol=1e-12 * np.abs(ref).max())


# ---------------------------------------------------------------------------------------------------------------------
# (b) linalg
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 5, 127, 128, 129, 300, 640, 1100])
def test_pdinv_against_lapack(n):
    rs = np.random.RandomState(n)
    B = rs.randn(n, n + 3)
    A = B @ B.T + 0.5 * n * np.eye(n)
    rc, Ai, L, Li, logdet = native.pdinv(A)
    assert rc == 0
    Lr = np.linalg.cholesky(A)
    assert_allclose(L, Lr, rtol=1e-10, atol=1e-12 * np.abs(Lr).max())
    assert_allclose(Li, np.linalg.inv(Lr), rtol=1e-9, atol=1e-12)
    assert_allclose(Ai, np.linalg.inv(A), rtol=1e-9, atol=1e-12 * np.abs(Ai).max())
    assert_allclose(logdet, 2 * np.log(np.diag(Lr)).sum(), rtol=1e-12)
    assert np.array_equal(Ai, Ai.T)
    assert np.all(np.triu(L, 1) == 0)


def test_pdinv_not_pd_reports_info():
    rs = np.random.RandomState(0)
    B = rs.randn(300, 300)
    A = B @ B.T
    w, V = np.linalg.eigh(A)
    w[0] = -1.0
    A = (V * w) @ V.T
    rc = native.pdinv(A)[0]
    assert rc > 0
    with pytest.raises(np.linalg.LinAlgError):
        native._lib.check(rc, "pdinv")


@pytest.mark.parametrize("n,nrhs", [(64, 1), (300, 3)])
def test_potrs(n, nrhs):
    rs = np.random.RandomState(n)
    B = rs.randn(n, n)
    A = B @ B.T + n * np.eye(n)
    L = np.linalg.cholesky(A)
    rhs = rs.randn(n, nrhs)
    assert_allclose(native.potrs(L, rhs), np.linalg.solve(A, rhs), rtol=1e-9, atol=1e-12)


# ---------------------------------------------------------------------------------------------------------------------
# (b)+(c) model: inference, gradients, posterior, predictions, acquisitions -- against the reference's own outputs
# ---------------------------------------------------------------------------------------------------------------------
def _fitted(g, cand_block=128):
    v, ls, n = _theta(g)
    m = native.NativeModel(g["kind"], g["ard"], g["X"].shape[1], 1, n_cap=g["X"].shape[0], cand_block=cand_block)
    m.set_data(g["X"], g["Y"])
    m.set_theta(v, ls, n)
    info, logL, grads = m.fit(True)
    assert info == 0
    return m, logL, grads


def test_inference_golden(golden):
    g = golden
    m, logL, grads = _fitted(g)
    ct = _cond_tol(g, 1.0)
    assert_allclose(logL, g["logL"], rtol=1e-9 * ct)
    assert_allclose(m.get("L"), g["L"], rtol=1e-9, atol=1e-12)
    assert_allclose(m.get("alpha"), g["alpha"], rtol=1e-9 * ct, atol=1e-9 * ct * np.abs(g["alpha"]).max())
    Wi = m.get("Wi")
    assert_allclose(Wi, g["Wi"], rtol=1e-9 * ct, atol=1e-9 * ct * np.abs(g["Wi"]).max())
    assert_allclose(m.get("dL_dK"), g["dL_dK"], rtol=1e-9 * ct, atol=1e-9 * ct * np.abs(g["dL_dK"]).max())
    assert_allclose(m.get("K"), g["K"], rtol=1e-9, atol=1e-300)
    assert_allclose(grads[0], g["grad_var"].ravel()[0], rtol=1e-7 * ct)
    assert_allclose(grads[1:-1], g["grad_len"], rtol=1e-7 * ct)
    assert_allclose(grads[-1], g["grad_noise"].ravel()[0], rtol=1e-7 * ct)
    m.close()


def test_predict_and_acquisition_golden(golden):
    g = golden
    m, _, _ = _fitted(g)
    ct = _cond_tol(g, 1.0)
    Xs = g["Xs"]
    mu, var = m.predict(Xs)
    assert_allclose(mu, g["pred_mu"], rtol=1e-9 * ct, atol=1e-11 * ct)
    assert_allclose(var, g["pred_var"], rtol=1e-9 * ct, atol=1e-13 * ct)
    _, var0 = m.predict(Xs, include_likelihood=False)
    assert_allclose(var0, g["pred_var_noiseless"], rtol=1e-8 * ct, atol=1e-12 * ct)
    mu2, cov = m.predict_full_cov(Xs)
    assert_allclose(mu2, g["pred_mu"], rtol=1e-9 * ct, atol=1e-11 * ct)
    assert_allclose(cov, g["pred_cov"], rtol=1e-9 * ct, atol=1e-11 * ct)
    dmu, dvar = m.predictive_gradients(Xs)
    assert_allclose(dmu, g["dmu_dX"], rtol=1e-7 * ct, atol=1e-9 * ct * np.abs(g["dmu_dX"]).max())
    assert_allclose(dvar, g["dv_dX"], rtol=1e-7 * ct, atol=1e-9 * ct * np.abs(g["dv_dX"]).max())
    fmin = m.fmin()
    assert_allclose(fmin, g["fmin"], rtol=1e-9 * ct)
    r = m.acquisition("EI", 0.01, fmin, Xs, with_gradients=True, want_moments=True)
    assert_allclose(r["m"], g["gpm_m"], rtol=1e-9 * ct, atol=1e-11 * ct)
    assert_allclose(r["s"], g["gpm_s"], rtol=1e-9 * ct)
    assert_allclose(r["dmdx"], g["gpm_dmdx"], rtol=1e-7 * ct, atol=1e-9 * ct * np.abs(g["gpm_dmdx"]).max())
    assert_allclose(r["dsdx"], g["gpm_dsdx"], rtol=1e-7 * ct, atol=1e-9 * ct * np.abs(g["gpm_dsdx"]).max())
    assert_allclose(r["f"], g["ei_g_f"], rtol=1e-8 * ct, atol=1e-14)
    assert_allclose(r["df"], g["ei_g_df"], rtol=1e-7 * ct, atol=1e-9 * ct * np.abs(g["ei_g_df"]).max())
    r0 = m.acquisition("EI", 0.01, fmin, Xs)
    assert_allclose(r0["f"], g["ei"], rtol=1e-8 * ct, atol=1e-14)
    rl = m.acquisition("LCB", 2.0, fmin, Xs, with_gradients=True)
    assert_allclose(rl["f"], g["lcb_g_f"], rtol=1e-9 * ct, atol=1e-11 * ct)
    assert_allclose(rl["df"], g["lcb_g_df"], rtol=1e-7 * ct, atol=1e-9 * ct * np.abs(g["lcb_g_df"]).max())
    # anchor selection: the k lowest scores, ties -> lowest index
    k = 5
    vals, idx, pts = m.acq_topk("EI", 0.01, fmin, Xs, k)
    order = np.argsort(g["ei"].ravel(), kind="stable")[:k]
    assert np.array_equal(idx, order)
    assert_allclose(vals, g["ei"].ravel()[order], rtol=1e-8 * ct, atol=1e-14)
    assert np.array_equal(pts, Xs[order])
    m.close()


# ---------------------------------------------------------------------------------------------------------------------
# seeded larger cases against the oracle (sizes the oracle finishes in seconds)
# ---------------------------------------------------------------------------------------------------------------------
def _synth(N, D, seed=1234):
    """SURVEY.md 8(d) generator."""
    rs = np.random.RandomState(seed)
    X = rs.uniform(0, 1, (N, D))
    w = rs.randn(D)
    Y = np.sin(X @ w)[:, None] + 0.05 * rs.randn(N, 1)
    Y = (Y - Y.mean()) / Y.std()
    ls = 0.5 + 0.5 * np.arange(D) / D
    return X, Y, ls


@pytest.mark.parametrize("kind,N,D", [("rbf", 1000, 8), ("mat52", 1500, 16), ("rbf", 2048, 3)])
def test_nll_grad_against_oracle(kind, N, D):
    X, Y, ls = _synth(N, D)
    logL, grads, post = O.log_likelihood_and_gradients(kind, X, Y, 1.0, ls, 1e-2, native=True)
    m = native.NativeModel(kind, True, D, 1, n_cap=N, cand_block=512)
    m.set_data(X, Y)
    m.set_theta(1.0, ls, 1e-2)
    info, l2, g2 = m.fit(True)
    assert info == 0
    assert_allclose(l2, logL, rtol=1e-9)
    assert_allclose(g2, grads, rtol=1e-7)
    # repeated evaluation is bitwise reproducible (fixed-order reductions; the L-BFGS-B trajectory depends on it)
    info, l3, g3 = m.fit(True)
    assert l3 == l2 and np.array_equal(g2, g3)
    # acquisition over more candidates than one block, with and without gradients
    st = O.GPState(kind, X, Y, 1.0, ls, 1e-2)
    Xc = np.random.RandomState(7).uniform(0, 1, (1200, D))
    fmin = m.fmin()
    assert_allclose(fmin, st.get_fmin(), rtol=1e-9)
    f_ref, df_ref = st.acquisition("EI", Xc, with_gradients=True, native=True)
    r = m.acquisition("EI", 0.01, fmin, Xc, with_gradients=True)
    assert_allclose(r["f"], f_ref, rtol=1e-7, atol=1e-12)
    assert_allclose(r["df"], df_ref, rtol=1e-6, atol=1e-9 * np.abs(df_ref).max())
    f_ref, df_ref = st.acquisition("LCB", Xc, with_gradients=True, native=True)
    r = m.acquisition("LCB", 2.0, fmin, Xc, with_gradients=True)
    assert_allclose(r["f"], f_ref, rtol=1e-9, atol=1e-11)
    assert_allclose(r["df"], df_ref, rtol=1e-7, atol=1e-9 * np.abs(df_ref).max())
    vals, idx, pts = m.acq_topk("LCB", 2.0, fmin, Xc, 5, index_offset=1000)
    order = np.argsort(f_ref.ravel(), kind="stable")[:5]
    assert np.array_equal(idx - 1000, order)
    m.close()


def test_model_grows_like_bo_loop():
    """set_XY with one more point per step (GPModel.updateModel, gpmodel.py:78-93): N crosses the 128 padding boundary."""
    X, Y, ls = _synth(140, 2)
    m = native.NativeModel("mat52", False, 2, 1, n_cap=256, cand_block=128)
    for n in (3, 5, 127, 128, 129, 140):
        m.set_data(X[:n], Y[:n])
        m.set_theta(1.3, [0.7], 1e-3)
        info, logL, grads = m.fit(True)
        assert info == 0
        l_ref, g_ref, _ = O.log_likelihood_and_gradients("mat52", X[:n], Y[:n], 1.3, [0.7], 1e-3, ard=False)
        assert_allclose(logL, l_ref, rtol=1e-9)
        assert_allclose(grads, g_ref, rtol=1e-7, atol=1e-9 * np.abs(g_ref).max())
    m.close()


def test_not_positive_definite_returns_info():
    """Duplicate inputs + zero noise: Ky is singular up to the 1e-8 jitter of exact_gaussian_inference.py:56."""
    X = np.zeros((200, 2))
    Y = np.ones((200, 1))
    m = native.NativeModel("rbf", True, 2, 1, n_cap=200, cand_block=128)
    m.set_data(X, Y)
    m.set_theta(1.0, [1.0, 1.0], 0.0)
    info, _, _ = m.fit(False, extra_jitter=-2e-8)   # net diagonal shift < 0 -> a non-positive pivot must be reported
    assert info > 0
    m.close()


@pytest.mark.parametrize("kind,N,D", [("rbf", 700, 6), ("mat52", 700, 6), ("mat52", 1301, 20), ("rbf", 333, 40), ("mat52", 90, 3)])
def test_small_candidate_counts_use_the_skinny_products(kind, N, D):
    """M = 1 .. 8 candidates (the L-BFGS-B refinement calls, optimizer.py:46-51) take the bandwidth-bound triangular
    matrix-vector path with the fused split-N row reductions; M = 9 takes the GEMM path.  Both must match the oracle and each
    other (every register-array size of the reduction kernel: D <= 4, 8, 16, 32, 64; N not a multiple of the chunk)."""
    X, Y, ls = _synth(N, D)
    st = O.GPState(kind, X, Y, 1.1, ls, 1e-3)
    m = native.NativeModel(kind, True, D, 1, n_cap=N, cand_block=256)
    m.set_data(X, Y)
    m.set_theta(1.1, ls, 1e-3)
    info, _, _ = m.fit(False)
    assert info == 0
    fmin = m.fmin()
    Xc = np.random.RandomState(3).uniform(0, 1, (9, D))
    f9 = m.acquisition("EI", 0.01, fmin, Xc, with_gradients=True, want_moments=True)
    for mc in (1, 2, 3, 4, 5, 8):
        f_ref, df_ref = st.acquisition("EI", Xc[:mc], with_gradients=True, native=True)
        r = m.acquisition("EI", 0.01, fmin, Xc[:mc], with_gradients=True, want_moments=True)
        assert_allclose(r["f"], f_ref, rtol=1e-7, atol=1e-12)
        assert_allclose(r["df"], df_ref, rtol=1e-6, atol=1e-9 * np.abs(df_ref).max())
        for key in ("f", "df", "m", "s", "dmdx", "dsdx"):
            assert_allclose(r[key], f9[key][:mc], rtol=1e-9, atol=1e-12 * max(1.0, np.abs(f9[key]).max()))
        mu, var = m.predict(Xc[:mc])
        mu_r, var_r = O.predict(kind, st.post, X, Xc[:mc], 1.1, ls, 1e-3)
        assert_allclose(mu, mu_r, rtol=1e-9, atol=1e-11)
        assert_allclose(var, var_r, rtol=1e-9, atol=1e-12)
        mu0, none = m.predict(Xc[:mc], want_var=False)                  # mean only
        assert none is None
        assert_allclose(mu0, mu, rtol=1e-11, atol=0)      # (the mean-only call sums the same products in another fixed order)
        dm, dv = m.predictive_gradients(Xc[:mc])
        dm9, dv9 = m.predictive_gradients(Xc)
        assert_allclose(dm, dm9[:mc], rtol=1e-9, atol=1e-12 * np.abs(dm9).max())
        assert_allclose(dv, dv9[:mc], rtol=1e-9, atol=1e-12 * np.abs(dv9).max())
        dm1, none = m.predictive_gradients(Xc[:mc], want_var=False)     # estimate_L's call: mean gradient only
        assert none is None
        assert_allclose(dm1, dm, rtol=1e-13, atol=0)
    m.close()


def test_acq_topk_full_single_pass():
    """Values, gradients and the k best from ONE pass equal the separate calls."""
    N, D = 900, 5
    X, Y, ls = _synth(N, D)
    m = native.NativeModel("mat52", True, D, 1, n_cap=N, cand_block=256)
    m.set_data(X, Y)
    m.set_theta(1.0, ls, 1e-3)
    assert m.fit(False)[0] == 0
    fmin = m.fmin()
    Xc = np.random.RandomState(8).uniform(0, 1, (700, D))
    ref = m.acquisition("EI", 0.01, fmin, Xc, with_gradients=True)
    v0, i0, p0 = m.acq_topk("EI", 0.01, fmin, Xc, 5, index_offset=100)
    vals, idx, pts, f, df = m.acq_topk_full("EI", 0.01, fmin, Xc, 5, index_offset=100)
    assert np.array_equal(f, ref["f"]) and np.array_equal(df, ref["df"])
    assert np.array_equal(idx, i0) and np.array_equal(vals, v0) and np.array_equal(pts, p0)
    m.close()


@pytest.mark.parametrize("kind,N,D,noise", [("rbf", 1500, 4, 1e-6), ("mat52", 2500, 6, 1e-6), ("rbf", 1200, 2, 1e-4)])
def test_ill_conditioned_exact_feval_models(kind, N, D, noise):
    """The exact_feval regime of GPyOpt (noise fixed at 1e-6, gpmodel.py:72-73) with smooth kernels: cond(Ky) is 1e8 .. 1e12.
    Two LAPACK builds only agree to cond * eps there; the explicit-inverse factorisation must not lose more than that."""
    X, Y, ls = _synth(N, D, seed=77)
    ls = ls * 0.6
    st = O.GPState(kind, X, Y, 1.0, ls, noise)
    w = np.linalg.eigvalsh(st.post.K + (noise + 1e-8) * np.eye(N))
    cond = w[-1] / w[0]
    tol = max(1.0, cond * 2.2e-16 / 1e-12)
    m = native.NativeModel(kind, True, D, 1, n_cap=N, cand_block=512)
    m.set_data(X, Y)
    m.set_theta(1.0, ls, noise)
    info, logL, g = m.fit(True)
    assert info == 0
    l_ref, g_ref, _ = O.log_likelihood_and_gradients(kind, X, Y, 1.0, ls, noise)
    assert_allclose(logL, l_ref, rtol=1e-9 * tol)
    assert_allclose(g, g_ref, rtol=1e-7 * tol, atol=1e-7 * tol * np.abs(g_ref).max())
    Xc = np.random.RandomState(2).uniform(0, 1, (300, D))
    mu, var = m.predict(Xc)
    mu_r, var_r = O.predict(kind, st.post, X, Xc, 1.0, ls, noise)
    assert_allclose(mu, mu_r, rtol=1e-9 * tol, atol=1e-10 * tol)
    assert_allclose(var, var_r, rtol=1e-9 * tol, atol=1e-12 * tol)
    assert np.all(var >= noise * 0.5)          # never below the noise floor by more than rounding
    print("cond(Ky) = %.2e, |dlogL|/|logL| = %.2e, max |dmu| = %.2e, max |dvar| = %.2e" %
          (cond, abs(logL - l_ref) / abs(l_ref), np.abs(mu - mu_r).max(), np.abs(var - var_r).max()))
    m.close()


@pytest.mark.parametrize("kind,n0,steps,grad_every", [("mat52", 100, (5, 23, 1, 150, 64), 2), ("rbf", 384, (128, 1, 300), 2),
                                                      ("mat52", 200, (56, 1, 127, 1, 300, 2), 1),
                                                      ("rbf", 130, (1, 1, 1, 130, 1), 3)])
def test_append_extends_the_factorisation(kind, n0, steps, grad_every):
    """gpb_model_append: the block rows of the new points are factorised against the resident factor; everything must equal a
    full refit on the extended data (within the padding block, across 128-boundaries, several blocks at once).  Ky^-1 is carried
    across appends as well (downdate of the recomputed block row + rank-r update), also when it is only asked for several appends
    later (grad_every > 1)."""
    D = 3
    X, Y, ls = _synth(n0 + sum(steps), D, seed=5)
    m = native.NativeModel(kind, True, D, 1, n_cap=X.shape[0], cand_block=128)
    ref = native.NativeModel(kind, True, D, 1, n_cap=X.shape[0], cand_block=128)
    theta = (1.2, ls, 1e-3)
    m.set_data(X[:n0], Y[:n0])
    m.set_theta(*theta)
    assert m.fit(False)[0] == 0
    n = n0
    Xc = np.random.RandomState(1).uniform(0, 1, (40, D))
    for i, b in enumerate(steps):
        Yn = Y[:n + b] * (1.0 + 0.01 * i)            # GPyOpt re-normalises all targets on every step
        want_grad = i % grad_every == 0
        info, logL, g = m.append(X[n:n + b], Yn, want_grad=want_grad)
        n += b
        assert info == 0 and m.n == n
        ref.set_data(X[:n], Yn)
        ref.set_theta(*theta)
        info, l_ref, g_ref = ref.fit(True)
        assert info == 0
        assert_allclose(logL, l_ref, rtol=1e-10)
        if want_grad:
            assert_allclose(g, g_ref, rtol=1e-7)
            if grad_every == 1:
                assert_allclose(m.get("Wi"), ref.get("Wi"), rtol=1e-7, atol=1e-9 * np.abs(ref.get("Wi")).max())
        for what in ("L", "Li", "alpha"):
            a, b_ = m.get(what), ref.get(what)
            assert_allclose(a, b_, rtol=1e-7, atol=1e-9 * np.abs(b_).max())
        mu, var = m.predict(Xc)
        mu_r, var_r = ref.predict(Xc)
        # different blocking of the same arithmetic: agreement to cond(Ky) * eps (the smooth RBF case has cond ~ 1e7)
        assert_allclose(mu, mu_r, rtol=1e-8, atol=1e-10)
        assert_allclose(var, var_r, rtol=1e-8, atol=1e-11)
        dm, dv = m.predictive_gradients(Xc)
        dm_r, dv_r = ref.predictive_gradients(Xc)
        assert_allclose(dm, dm_r, rtol=1e-7, atol=1e-9 * np.abs(dm_r).max())
        assert_allclose(dv, dv_r, rtol=1e-7, atol=1e-9 * np.abs(dv_r).max())
    assert_allclose(m.get("Wi"), ref.get("Wi"), rtol=1e-7, atol=1e-9 * np.abs(ref.get("Wi")).max())
    m.close()
    ref.close()


def test_edge_sizes_one_point_empty_candidates_two_outputs():
    """N = 1 and N = 2 training points, an empty candidate set, and P = 2 output columns (exact_gaussian_inference.py:62,70 carry
    the factor P; posterior.py:276 gives one mean column per output)."""
    rs = np.random.RandomState(9)
    for n in (1, 2):
        X, Y = rs.uniform(0, 1, (n, 3)), rs.randn(n, 1)
        ls = np.array([0.4, 0.6, 0.9])
        m = native.NativeModel("rbf", True, 3, 1, n_cap=4, cand_block=128)
        m.set_data(X, Y)
        m.set_theta(0.8, ls, 0.1)
        info, logL, g = m.fit(True)
        l_ref, g_ref, post = O.log_likelihood_and_gradients("rbf", X, Y, 0.8, ls, 0.1)
        assert info == 0
        assert_allclose(logL, l_ref, rtol=1e-12)
        assert_allclose(g, g_ref, rtol=1e-9, atol=1e-14)
        Xc = rs.uniform(0, 1, (5, 3))
        mu, var = m.predict(Xc)
        mu_r, var_r = O.predict("rbf", post, X, Xc, 0.8, ls, 0.1)
        assert_allclose(mu, mu_r, rtol=1e-12, atol=1e-15)
        assert_allclose(var, var_r, rtol=1e-12)
        mu0, var0 = m.predict(np.zeros((0, 3)))
        assert mu0.shape == (0, 1) and var0.shape == (0, 1)
        r = m.acquisition("EI", 0.01, m.fmin(), np.zeros((0, 3)), with_gradients=True)
        assert r["f"].shape == (0, 1) and r["df"].shape == (0, 3)
        m.close()
    # two output columns
    N, D, P = 200, 4, 2
    X = rs.uniform(0, 1, (N, D))
    Y = np.stack([np.sin(3 * X[:, 0]) + X[:, 1], np.cos(2 * X[:, 2]) * X[:, 3]], 1) + 0.05 * rs.randn(N, P)
    ls = np.array([0.5, 0.7, 0.9, 1.1])
    for kind in ("rbf", "mat52"):
        m = native.NativeModel(kind, True, D, P, n_cap=N, cand_block=128)
        m.set_data(X, Y)
        m.set_theta(1.4, ls, 0.02)
        info, logL, g = m.fit(True)
        assert info == 0
        l_ref, g_ref, post = O.log_likelihood_and_gradients(kind, X, Y, 1.4, ls, 0.02)
        assert_allclose(logL, l_ref, rtol=1e-10)
        assert_allclose(g, g_ref, rtol=1e-7, atol=1e-10)
        assert_allclose(m.get("alpha"), post.woodbury_vector, rtol=1e-8, atol=1e-10)
        for mc in (3, 40):                       # the skinny path is single-output: both sizes take the general one
            Xc = rs.uniform(0, 1, (mc, D))
            mu, var = m.predict(Xc)
            mu_r, var_r = O.predict(kind, post, X, Xc, 1.4, ls, 0.02)
            assert mu.shape == (mc, P)
            assert_allclose(mu, mu_r, rtol=1e-9, atol=1e-11)
            assert_allclose(var, var_r, rtol=1e-9, atol=1e-12)
        m.close()


@pytest.mark.parametrize("kind", ["rbf", "mat52"])
def test_wide_inputs_d48_block_and_skinny_routes_match_the_oracle(kind):
    """32 < D <= 64: the gradient kernels take their two-sweep form (no register spills); fit, M = 3 and M = 200 acquisition calls."""
    rs = np.random.RandomState(48)
    n, d = 300, 48
    X = rs.uniform(0, 1, (n, d))
    Y = np.sin(X[:, :5].sum(1))[:, None] + 0.05 * rs.randn(n, 1)
    Y = (Y - Y.mean()) / Y.std()
    ls = 1.5 + 0.02 * np.arange(d)
    m = native.NativeModel(kind, True, d, 1, n_cap=n, cand_block=256)
    m.set_data(X, Y)
    m.set_theta(1.2, ls, 1e-2)
    info, logL, g = m.fit(True)
    assert info == 0
    l_ref, g_ref, _ = O.log_likelihood_and_gradients(kind, X, Y, 1.2, ls, 1e-2)
    assert_allclose(logL, l_ref, rtol=1e-9)
    assert_allclose(g, g_ref, rtol=1e-7, atol=1e-9 * np.abs(g_ref).max())
    st = O.GPState(kind, X, Y, 1.2, ls, 1e-2)
    fmin = m.fmin()
    for mc in (3, 200):
        Xc = rs.uniform(0, 1, (mc, d))
        r = m.acquisition("EI", 0.01, fmin, Xc, with_gradients=True)
        f_ref, df_ref = st.acquisition("EI", Xc, with_gradients=True)
        assert_allclose(r["f"], f_ref, rtol=1e-7, atol=1e-12)
        assert_allclose(r["df"], df_ref, rtol=1e-6, atol=1e-9 * np.abs(df_ref).max())
    G = rs.randn(40, n)
    gx = native.kern_gradients_X(kind, G, X[:40] + 0.01, X, 1.2, ls)
    gx_ref = O.gradients_X(kind, G, X[:40] + 0.01, X, 1.2, ls)
    assert_allclose(gx, gx_ref, rtol=1e-7, atol=1e-9 * np.abs(gx_ref).max())
    m.close()


def test_cooperative_block_solver_is_bit_identical_to_the_launch_chain():
    """GPB_COOP_N selects the persistent cooperative kernel for diagonal blocks of the recursion (measured slower, off by default:
    profiles/r2q_coop_solver.json).  It is read once per process, so the comparison runs in two child processes."""
    import subprocess
    import sys
    code = ("import sys, numpy as np; sys.path.insert(0, '.');"
            "from gaussian_process_optimization_b200 import native;"
            "rs = np.random.RandomState(3); n = 900; B = rs.randn(n, n + 2); A = B @ B.T + 0.5 * n * np.eye(n);"
            "rc, Ai, L, Li, ld = native.pdinv(A); assert rc == 0;"
            "import hashlib; print(hashlib.sha256(L.tobytes() + Li.tobytes() + Ai.tobytes()).hexdigest(), repr(ld))")
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for coop in ("0", "1024"):
        env = dict(os.environ, GPB_COOP_N=coop)
        r = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(r.stdout.strip().splitlines()[-1])
    assert outs[0] == outs[1]
