"""Parity at BASELINE.json's full sizes through size-independent properties (the oracle needs minutes at N = 16384).

Headline configuration: N = 16384, D = 16, synthetic data of SURVEY.md 8(d).  Checked on the device (cuBLAS fp64 products
via torch are the independent reference for the residuals):
  factorisation   L L^T = Ky,  M L = I,  Ky^-1 Ky = I,  Ky alpha = Y,  log-likelihood recomputed from L and alpha
  gradients       central finite differences of the log-likelihood in every parameter class (variance, lengthscale, noise)
  prediction      at training inputs  mu = Y - s alpha  and  var = s - s^2 (Ky^-1)_ii  with s = noise + 1e-8 (closed forms)
  acquisition     top-k == stable argsort of the scores; sharded ranges merged == unsharded; bitwise repeatability
"""
import numpy as np
import pytest
from numpy.testing import assert_allclose

pytestmark = pytest.mark.gpu

native = pytest.importorskip("gaussian_process_optimization_b200.native")
from gaussian_process_optimization_b200 import sharded  # noqa: E402

N, D = 16384, 16


def _synth(n, d, seed=1234):
    rs = np.random.RandomState(seed)
    X = rs.uniform(0, 1, (n, d))
    w = rs.randn(d)
    Y = np.sin(X @ w)[:, None] + 0.05 * rs.randn(n, 1)
    Y = (Y - Y.mean()) / Y.std()
    return X, Y, 0.5 + 0.5 * np.arange(d) / d


@pytest.fixture(scope="module")
def fitted():
    import torch
    X, Y, ls = _synth(N, D)
    m = native.NativeModel("mat52", True, D, 1, n_cap=N, cand_block=2048)
    m.set_data(X, Y)
    theta = (1.3, ls, 1e-2)
    m.set_theta(*theta)
    info, logL, g = m.fit(True)
    assert info == 0
    yield m, X, Y, theta, logL, g, torch
    m.close()


def test_factorisation_residuals(fitted):
    m, X, Y, (v, ls, noise), logL, g, torch = fitted
    dev = torch.device("cuda")
    new = lambda: torch.empty((N, N), dtype=torch.float64, device=dev)  # noqa: E731
    K = m.get("K", out=new())
    Ky = K + (noise + 1e-8) * torch.eye(N, dtype=torch.float64, device=dev)
    del K
    L = m.get("L", out=new())
    R = L @ L.T
    R -= Ky
    assert float(R.abs().max()) <= 1e-12 * float(Ky.abs().max()) * N ** 0.5
    Li = m.get("Li", out=R)
    P = Li @ L
    P -= torch.eye(N, dtype=torch.float64, device=dev)
    assert float(P.abs().max()) <= 1e-9
    del P, Li, R
    Wi = m.get("Wi", out=new())
    assert float((Wi - Wi.T).abs().max()) == 0.0          # symmetrified like linalg.py:144
    Q = Wi @ Ky
    Q -= torch.eye(N, dtype=torch.float64, device=dev)
    assert float(Q.abs().max()) <= 1e-8
    del Q
    alpha = torch.from_numpy(m.get("alpha")).to(dev)
    Yd = torch.from_numpy(Y).to(dev)
    assert float((Ky @ alpha - Yd).abs().max()) <= 1e-9
    logdet = 2.0 * float(torch.log(torch.diagonal(L)).sum())
    ll = 0.5 * (-N * np.log(2 * np.pi) - logdet - float((alpha * Yd).sum()))
    assert_allclose(logL, ll, rtol=1e-12)
    # dL/dnoise = tr(dL_dK) = 0.5 (alpha^T alpha - tr(Ky^-1))      exact_gaussian_inference.py:70-72
    assert_allclose(g[-1], 0.5 * (float((alpha * alpha).sum()) - float(torch.diagonal(Wi).sum())), rtol=1e-10)
    # closed forms at the training inputs
    s = noise + 1e-8
    idx = np.arange(0, N, 37)[:300]
    mu, var = m.predict(X[idx], include_likelihood=False)
    a_np = alpha.cpu().numpy()
    assert_allclose(mu, Y[idx] - s * a_np[idx], rtol=1e-9, atol=1e-10)
    wii = torch.diagonal(Wi).cpu().numpy()[idx]
    assert_allclose(var.ravel(), s - s * s * wii, rtol=1e-6, atol=1e-11)


def test_gradients_against_finite_differences(fitted):
    m, X, Y, (v, ls, noise), logL, g, torch = fitted

    def ll(v_, ls_, nz_):
        m.set_theta(v_, ls_, nz_)
        info, val, _ = m.fit(False)
        assert info == 0
        return val

    h = 1e-5
    fd_v = (ll(v * (1 + h), ls, noise) - ll(v * (1 - h), ls, noise)) / (2 * v * h)
    assert_allclose(g[0], fd_v, rtol=1e-5, atol=1e-4)
    fd_n = (ll(v, ls, noise * (1 + h)) - ll(v, ls, noise * (1 - h))) / (2 * noise * h)
    assert_allclose(g[-1], fd_n, rtol=1e-5, atol=1e-4)
    for q in (0, D - 1):
        lp, lm = ls.copy(), ls.copy()
        lp[q] *= 1 + h
        lm[q] *= 1 - h
        fd_l = (ll(v, lp, noise) - ll(v, lm, noise)) / (2 * ls[q] * h)
        assert_allclose(g[1 + q], fd_l, rtol=1e-5, atol=1e-4)
    # restore the fitted state for the tests below and check bitwise repeatability of the whole evaluation
    m.set_theta(v, ls, noise)
    info, logL2, g2 = m.fit(True)
    assert info == 0 and logL2 == logL and np.array_equal(g, g2)


def test_acquisition_topk_properties(fitted):
    m, X, Y, theta, logL, g, torch = fitted
    fmin = m.fmin()
    assert_allclose(fmin, (Y - (theta[2] + 1e-8) * m.get("alpha")).min(), rtol=1e-9)     # min of the closed-form mean
    Xc = np.random.RandomState(4321).uniform(0, 1, (2 ** 14, D))
    f = m.acquisition("EI", 0.01, fmin, Xc)["f"].ravel()
    assert np.all(f <= 0.0) and np.all(np.isfinite(f))
    vals, idx, pts = m.acq_topk("EI", 0.01, fmin, Xc, 5)
    order = np.argsort(f, kind="stable")[:5]
    assert np.array_equal(idx, order) and np.array_equal(vals, f[order]) and np.array_equal(pts, Xc[order])
    # sharding the candidate set and merging the per-shard top-5 gives the same anchors (world sizes 2, 4, 8)
    for world in (2, 4, 8):
        parts = []
        for r in range(world):
            lo, hi = sharded.divide_candidates(Xc.shape[0], r, world)
            parts.append(m.acq_topk("EI", 0.01, fmin, Xc[lo:hi], 5, index_offset=lo))
        mv, mi, mp = sharded.merge_topk(np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts]),
                                        np.concatenate([p[2] for p in parts]), 5)
        assert np.array_equal(mi, order) and np.array_equal(mv, f[order]) and np.array_equal(mp, Xc[order])
    # value + gradient pass returns the same values as the value-only pass, and the gradient matches a finite difference
    r = m.acquisition("EI", 0.01, fmin, Xc[:64], with_gradients=True)
    assert_allclose(r["f"].ravel(), f[:64], rtol=1e-12, atol=1e-300)
    c = int(np.argmin(f[:64]))
    e = np.zeros(D)
    e[3] = 1e-6
    fd = (m.acquisition("EI", 0.01, fmin, (Xc[c] + e)[None])["f"][0, 0] - m.acquisition("EI", 0.01, fmin, (Xc[c] - e)[None])["f"][0, 0]) / 2e-6
    assert_allclose(r["df"][c, 3], fd, rtol=1e-4, atol=1e-9)


def test_int8_engine_at_the_headline_size(fitted):
    """The experimental int8 tensor-core engine (csrc/gpb_ozaki.cu, off by default) on the two top recursion levels, Ky^-1 = M^T M and
    the predictive products: same log-likelihood, gradients, residuals and anchors as the fp64 engine at N = 16384."""
    m, X, Y, (v, ls, noise), logL, g, torch = fitted
    dev = torch.device("cuda")
    fmin0 = m.fmin()
    Xc = np.random.RandomState(99).uniform(0, 1, (2 ** 13, D))
    vals0, idx0, pts0, f0, _ = m.acq_topk_full("EI", 0.01, fmin0, Xc, 5, with_gradients=False)
    try:
        # (18 and 16: the engine's modular mode with that many moduli -- 62 and 56 bits per operand, one int8 product per modulus)
        for digits, tl, tg in ((18, 1e-12, 1e-10), (16, 1e-10, 1e-8), (8, 1e-12, 1e-10), (7, 1e-10, 1e-8)):
            native.set_ozaki(8192, digits)
            m.set_theta(v, ls, noise)
            info, l1, g1 = m.fit(True)
            assert info == 0
            assert_allclose(l1, logL, rtol=tl)
            assert_allclose(g1, g, rtol=tg, atol=tg * np.abs(g).max())
        # residual of the inverse built by the engine (7 digits, the looser setting): Ky^-1 Ky = I
        K = m.get("K", out=torch.empty((N, N), dtype=torch.float64, device=dev))
        K += (noise + 1e-8) * torch.eye(N, dtype=torch.float64, device=dev)
        Wi = m.get("Wi", out=torch.empty((N, N), dtype=torch.float64, device=dev))
        Q = Wi @ K
        Q -= torch.eye(N, dtype=torch.float64, device=dev)
        assert float(Q.abs().max()) <= 1e-8
        del Q, K, Wi
        fmin1 = m.fmin()
        assert_allclose(fmin1, fmin0, rtol=1e-9)
        vals1, idx1, pts1, f1, _ = m.acq_topk_full("EI", 0.01, fmin1, Xc, 5, with_gradients=False)
        assert np.array_equal(idx1, idx0)
        assert_allclose(f1, f0, rtol=1e-7, atol=1e-9 * np.abs(f0).max())
    finally:
        native.set_ozaki(0)
        m.set_theta(v, ls, noise)
        m.fit(True)                                   # leave the shared model as the fp64 engine built it


def test_n32768_d20_end_state_of_config5():
    """BASELINE.json config 5 grows the model to N = 32768 (D = 20): three 8.6 GB matrices resident on one GPU.  Checked through
    the closed forms at the training inputs (no second N x N copy needed) and one finite difference."""
    n, d = 32768, 20
    X, Y, ls = _synth(n, d, seed=20)
    m = native.NativeModel("mat52", True, d, 1, n_cap=n, cand_block=1024)
    m.set_data(X, Y)
    v, noise = 1.1, 1e-2
    m.set_theta(v, ls, noise)
    info, logL, g = m.fit(True)
    assert info == 0 and np.isfinite(logL) and np.all(np.isfinite(g))
    s = noise + 1e-8
    alpha = m.get("alpha")
    idx = np.arange(0, n, 113)[:256]
    mu, var = m.predict(X[idx], include_likelihood=False)
    assert_allclose(mu, Y[idx] - s * alpha[idx], rtol=1e-9, atol=1e-10)
    assert np.all(var > 0) and np.all(var < s)                      # var = s - s^2 (Ky^-1)_ii, and (Ky^-1)_ii > 0
    h = 1e-5
    m.set_theta(v, ls, noise * (1 + h))
    _, lp, _ = m.fit(False)
    m.set_theta(v, ls, noise * (1 - h))
    _, lm, _ = m.fit(False)
    assert_allclose(g[-1], (lp - lm) / (2 * noise * h), rtol=1e-5, atol=1e-4)
    m.close()
