"""Parity at BASELINE.json's OWN sizes against vectors produced by the reference's own sources (tests/golden/fullsize/*.npz,
generator tests/golden/make_golden_fullsize.py: GPy GPRegression / GPyOpt GPModel / AcquisitionEI / AcquisitionLCB executed
unmodified in the build container at N = 4096 (config 2) and N = 16384 (headline metric, configs 3 and 4)).

Through the C ABI (native.NativeModel), at north_star's bars: rtol 1e-9 on the log-likelihood and the predictions, 1e-7 on
gradients, identical top-5 candidates.  Quantities whose conditioning is cond(Ky) * eps (alpha, Ky^-1, the predictive variance
near the data and what is built from them) carry that factor explicitly, taken from the bound stored in the fixture
(||Ky||_inf ||Ky^-1||_inf) -- two LAPACK builds do not agree better than that either (the fixture records how far the CPU oracle
is from the reference on the same inputs: `oracle_vs_ref_*`).

Every case also runs with the experimental int8 tensor-core engine forced on (modular mode, 18 moduli, every product of >= 256
rows on tcgen05) at the SAME bars.
"""
import os

import numpy as np
import pytest
from numpy.testing import assert_allclose

from conftest import GOLDEN_DIR

pytestmark = pytest.mark.gpu

native = pytest.importorskip("gaussian_process_optimization_b200.native")

FULL_DIR = os.path.join(GOLDEN_DIR, "fullsize")
CASES = sorted(f[:-4] for f in os.listdir(FULL_DIR) if f.endswith(".npz")) if os.path.isdir(FULL_DIR) else []
ENGINES = ["fp64_dmma", "int8_crt18"]


def _synth(n, d, seed):
    rs = np.random.RandomState(seed)
    X = rs.uniform(0, 1, (n, d))
    w = rs.randn(d)
    Y = np.sin(X @ w)[:, None] + 0.05 * rs.randn(n, 1)
    Y = (Y - Y.mean()) / Y.std()
    return X, Y


def _load(name):
    z = np.load(os.path.join(FULL_DIR, name + ".npz"), allow_pickle=False)
    g = {k: z[k] for k in z.files}
    for k in ("kind",):
        g[k] = str(g[k])
    for k in ("N", "D", "data_seed", "cand_seed"):
        g[k] = int(g[k])
    for k in ("variance", "noise", "logL", "fmin", "cond_bound", "Wi_trace"):
        g[k] = float(g[k])
    return g


def _cond_factor(g):
    """max(1, cond(Ky) eps / 1e-12) with the fixture's bound on cond(Ky): 1 for well-conditioned models."""
    return max(1.0, g["cond_bound"] * 2.2e-16 / 1e-12)


@pytest.fixture(scope="module", params=[(c, e) for c in CASES for e in ENGINES], ids=lambda p: "%s-%s" % p)
def fitted(request):
    name, engine = request.param
    g = _load(name)
    X, Y = _synth(g["N"], g["D"], g["data_seed"])
    if engine == "int8_crt18":
        native.set_ozaki(256, 18)
    try:
        m = native.NativeModel(g["kind"], True, g["D"], 1, n_cap=g["N"], cand_block=2048)
        m.set_data(X, Y)
        m.set_theta(g["variance"], g["lengthscale"], g["noise"])
        info, logL, grads = m.fit(True)
        assert info == 0
        Xc = np.random.RandomState(g["cand_seed"]).uniform(0, 1, (2 ** 16, g["D"]))
        yield g, m, logL, grads, Xc, engine
        m.close()
    finally:
        native.set_ozaki(0)


def test_log_likelihood_and_gradients(fitted):
    g, m, logL, grads, Xc, engine = fitted
    assert_allclose(logL, g["logL"], rtol=1e-9)
    assert_allclose(grads, g["grads"], rtol=1e-7, atol=1e-7 * np.abs(g["grads"]).max())


def test_posterior_state(fitted):
    g, m, logL, grads, Xc, engine = fitted
    rows, cf = g["rows"], _cond_factor(g)
    import torch
    n = g["N"]
    alpha = m.get("alpha")[rows, 0]
    assert_allclose(alpha, g["alpha_rows"], rtol=1e-9 * cf, atol=1e-9 * cf * np.abs(g["alpha_rows"]).max())
    L = m.get("L", out=torch.empty((n, n), dtype=torch.float64, device="cuda"))
    assert_allclose(torch.diagonal(L)[torch.from_numpy(rows).cuda()].cpu().numpy(), g["L_diag_rows"], rtol=1e-9 * cf)
    last = L[n - 1].cpu().numpy()[rows]
    assert_allclose(last, g["L_lastrow_rows"], rtol=1e-9 * cf, atol=1e-9 * cf * np.abs(g["L_lastrow_rows"]).max())
    del L
    Wi = m.get("Wi", out=torch.empty((n, n), dtype=torch.float64, device="cuda"))
    wd = torch.diagonal(Wi)
    assert_allclose(wd[torch.from_numpy(rows).cuda()].cpu().numpy(), g["Wi_diag_rows"], rtol=1e-9 * cf)
    assert_allclose(float(wd.sum()), g["Wi_trace"], rtol=1e-9 * cf)
    del Wi


def test_fmin_and_moments(fitted):
    g, m, logL, grads, Xc, engine = fitted
    cf = _cond_factor(g)
    fmin = m.fmin()
    assert_allclose(fmin, g["fmin"], rtol=1e-9 * cf)
    Xg = Xc[:g["gpm_m"].size]
    r = m.acquisition("EI", 0.01, fmin, Xg, with_gradients=True, want_moments=True)
    assert_allclose(r["m"].ravel(), g["gpm_m"], rtol=1e-9 * cf, atol=1e-9 * cf * np.abs(g["gpm_m"]).max())
    assert_allclose(r["s"].ravel(), g["gpm_s"], rtol=1e-9 * cf)
    assert_allclose(r["dmdx"], g["gpm_dmdx"], rtol=1e-7, atol=1e-7 * cf * np.abs(g["gpm_dmdx"]).max())
    assert_allclose(r["dsdx"], g["gpm_dsdx"], rtol=1e-7, atol=1e-7 * cf * np.abs(g["gpm_dsdx"]).max())


def test_ei_values_gradients_and_top5(fitted):
    g, m, logL, grads, Xc, engine = fitted
    cf = _cond_factor(g)
    fmin = m.fmin()
    vals, idx, pts, f, df = m.acq_topk_full("EI", 0.01, fmin, Xc, 5, with_gradients=False)
    assert np.array_equal(idx, g["ei_top5_idx"])                                   # identical argmax candidates
    assert np.array_equal(pts, Xc[g["ei_top5_idx"]])
    assert_allclose(vals, g["ei_top5_val"], rtol=1e-7 * cf)
    scale = np.abs(g["ei_f"]).max()
    assert_allclose(f.ravel(), g["ei_f"], rtol=1e-7 * cf, atol=1e-9 * cf * scale)
    Xg = Xc[:g["ei_g_f"].size]
    r = m.acquisition("EI", 0.01, fmin, Xg, with_gradients=True)
    assert_allclose(r["f"].ravel(), g["ei_g_f"], rtol=1e-7 * cf, atol=1e-9 * cf * scale)
    assert_allclose(r["df"], g["ei_g_df"], rtol=1e-7 * cf, atol=1e-7 * cf * np.abs(g["ei_g_df"]).max())


def test_lcb_values_gradients_and_top5(fitted):
    g, m, logL, grads, Xc, engine = fitted
    cf = _cond_factor(g)
    Xl = Xc[:g["lcb_f"].size]
    vals, idx, pts, f, df = m.acq_topk_full("LCB", 2.0, 0.0, Xl, 5, with_gradients=False)
    assert np.array_equal(idx, g["lcb_top5_idx"])
    assert_allclose(f.ravel(), g["lcb_f"], rtol=1e-9 * cf, atol=1e-9 * cf * np.abs(g["lcb_f"]).max())
    Xg = Xc[:g["lcb_g_f"].size]
    r = m.acquisition("LCB", 2.0, 0.0, Xg, with_gradients=True)
    assert_allclose(r["f"].ravel(), g["lcb_g_f"], rtol=1e-9 * cf, atol=1e-9 * cf * np.abs(g["lcb_g_f"]).max())
    assert_allclose(r["df"], g["lcb_g_df"], rtol=1e-7 * cf, atol=1e-7 * cf * np.abs(g["lcb_g_df"]).max())
