"""Generate tests/golden/fullsize/*.npz: golden vectors at BASELINE.json's OWN sizes, produced by the reference's own sources.

Container-only (needs /root/reference; minutes of CPU per case, ~25 GB of host memory at N = 16384):

    python tests/golden/make_golden_fullsize.py [case ...]

Producer = the UNMODIFIED reference numerics executed through tests/golden/ref_harness.py (GPy GPRegression ->
ExactGaussianInference.inference -> Stationary.update_gradients_full; GPyOpt GPModel.predict(_withGradients) / get_fmin;
AcquisitionEI / AcquisitionLCB .acquisition_function(_withGradients)).  The CPU oracle (oracle/gp_oracle.py) is run on the same
inputs beside it and its agreement with the reference is stored in the fixture (`oracle_vs_ref_*`): that pins the oracle at the
sizes the small fixtures of make_golden.py cannot reach.

Inputs are NOT stored (they are regenerated from the seeds by the tests): training set = SURVEY.md 8(d) generator
`RandomState(1234)`, candidates = the first rows of chunk 0 = `RandomState(4321).uniform(0, 1, (2**20, D))`.

Cases (BASELINE.json configs; theta of SURVEY 8(d): variance 1, lengthscale_q = 0.5 + 0.5 q / D):
  config2_rbf_ard_n4096_d8            noise 1e-2      configs[1]
  headline_rbf_ard_n16384_d16         noise 1e-2      the metric's configuration
  config3_mat52_ard_n16384_d16        noise 1e-2      configs[2] (starting point of the optimize run)
  config4_mat52_ard_n16384_d16_exact  noise 1e-6      configs[3] (the exact_feval noise level, gpmodel.py:72-73)
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import ref_harness as rh  # noqa: E402

OUT = os.path.join(HERE, "fullsize")

CASES = {
    # name: (kind, N, D, noise, rows scored by EI (value only), candidate rows per reference call)
    "config2_rbf_ard_n4096_d8": ("rbf", 4096, 8, 1e-2, 2 ** 16, 2 ** 14),
    "headline_rbf_ard_n16384_d16": ("rbf", 16384, 16, 1e-2, 2 ** 16, 2 ** 14),
    "config3_mat52_ard_n16384_d16": ("mat52", 16384, 16, 1e-2, 2 ** 16, 2 ** 14),
    "config4_mat52_ard_n16384_d16_exact": ("mat52", 16384, 16, 1e-6, 2 ** 16, 2 ** 14),
}
M_GRAD = 2 ** 11        # rows with value + gradient (EI and LCB), posterior moments
M_LCB = 2 ** 14         # rows scored by LCB (value only)


def synth(N, D, seed=1234):
    """SURVEY.md 8(d) (identical to bench.py: synth)."""
    rs = np.random.RandomState(seed)
    X = rs.uniform(0, 1, (N, D))
    w = rs.randn(D)
    Y = np.sin(X @ w)[:, None] + 0.05 * rs.randn(N, 1)
    Y = (Y - Y.mean()) / Y.std()
    return X, Y, 0.5 + 0.5 * np.arange(D) / D


def candidates(D, rows):
    """First `rows` rows of chunk 0 of SURVEY.md 8(d)'s candidate set (row-major draw order -> a prefix of the stream)."""
    return np.random.RandomState(4321).uniform(0, 1, (rows, D))


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def run_case(ns, name):
    from oracle import gp_oracle as O
    kind, N, D, noise, m_ei, per_call = CASES[name]
    X, Y, ls = synth(N, D)
    variance = 1.0
    t0 = time.time()
    m = rh.make_model(ns, kind, X, Y, variance, ls, noise, ard=True)      # runs the reference's GP.parameters_changed
    k = m.kern
    out = dict(kind=kind, N=N, D=D, variance=variance, lengthscale=ls, noise=noise, data_seed=1234, cand_seed=4321)
    out["logL"] = float(m.log_likelihood())
    out["grads"] = np.concatenate([np.ravel(k.variance.gradient), np.ravel(k.lengthscale.gradient),
                                   np.ravel(m.likelihood.variance.gradient)])
    rows = np.arange(0, N, N // 256)
    out["rows"] = rows
    out["alpha_rows"] = np.array(m.posterior.woodbury_vector)[rows, 0]
    L = np.asarray(m.posterior.woodbury_chol)
    out["L_diag_rows"] = np.diag(L)[rows].copy()
    out["L_lastrow_rows"] = L[N - 1, rows].copy()
    Wi = np.asarray(m.posterior.woodbury_inv)
    out["Wi_diag_rows"] = np.diag(Wi)[rows].copy()
    out["Wi_trace"] = float(np.trace(Wi))
    print("  [%s] reference inference %.0f s  logL=%.12g" % (name, time.time() - t0, out["logL"]), flush=True)

    # oracle on the same inputs (also the source of the condition estimate the tests scale their tolerances with)
    t0 = time.time()
    lo, go, post = O.log_likelihood_and_gradients(kind, X, Y, variance, ls, noise, native=O.ref_native() is not None)
    out["oracle_vs_ref_logL"] = abs(lo - out["logL"]) / abs(out["logL"])
    out["oracle_vs_ref_grads"] = rel(go, out["grads"])
    out["oracle_vs_ref_alpha"] = rel(post.woodbury_vector[rows, 0], out["alpha_rows"])
    # cond(Ky) estimate: lambda_max <= ||Ky||_inf, lambda_min >= 1 / ||Ky^-1||_inf  (both matrices are at hand)
    Ky_inf = float(np.max(np.sum(np.abs(post.K), axis=1)) + noise + 1e-8)
    Wi_inf = float(np.max(np.sum(np.abs(post.woodbury_inv), axis=1)))
    out["cond_bound"] = Ky_inf * Wi_inf
    print("  [%s] oracle %.0f s  |dlogL|=%.2e |dgrad|=%.2e cond<=%.2e" % (name, time.time() - t0, out["oracle_vs_ref_logL"],
                                                                         out["oracle_vs_ref_grads"], out["cond_bound"]), flush=True)
    del post, L, Wi

    gm = rh.make_gpmodel(ns, m)
    sp = rh._Space()
    ei = ns.AcquisitionEI(gm, sp, optimizer=None, jitter=0.01)
    lcb = ns.AcquisitionLCB(gm, sp, optimizer=None, exploration_weight=2)
    t0 = time.time()
    out["fmin"] = float(gm.get_fmin())
    Xc = candidates(D, m_ei)
    # value only, all m_ei rows (the reference recomputes get_fmin inside every call, gpmodel.py:125-129 / EI.py:36)
    f = np.empty(m_ei)
    for a in range(0, m_ei, per_call):
        f[a:a + per_call] = np.asarray(ei.acquisition_function(Xc[a:a + per_call])).ravel()
        print("  [%s] EI rows %d..%d  %.0f s" % (name, a, a + per_call, time.time() - t0), flush=True)
    out["ei_f"] = f
    order = np.argsort(f, kind="stable")[:5]                  # anchor_points_generator.py:58-63 (np.argsort, 5 lowest)
    out["ei_top5_idx"] = order.astype(np.int64)
    out["ei_top5_val"] = f[order]
    out["lcb_f"] = np.asarray(lcb.acquisition_function(Xc[:M_LCB])).ravel()
    o2 = np.argsort(out["lcb_f"], kind="stable")[:5]
    out["lcb_top5_idx"] = o2.astype(np.int64)
    # value + gradient and the posterior moments behind them on the first M_GRAD rows
    Xg = Xc[:M_GRAD]
    mm, ss, dmdx, dsdx = gm.predict_withGradients(Xg)
    out["gpm_m"], out["gpm_s"] = np.asarray(mm).ravel(), np.asarray(ss).ravel()
    out["gpm_dmdx"], out["gpm_dsdx"] = np.asarray(dmdx), np.asarray(dsdx)
    fe, dfe = ei.acquisition_function_withGradients(Xg)
    out["ei_g_f"], out["ei_g_df"] = np.asarray(fe).ravel(), np.asarray(dfe)
    fl, dfl = lcb.acquisition_function_withGradients(Xg)
    out["lcb_g_f"], out["lcb_g_df"] = np.asarray(fl).ravel(), np.asarray(dfl)
    print("  [%s] reference acquisition %.0f s  fmin=%.9g top5=%s" % (name, time.time() - t0, out["fmin"], order.tolist()), flush=True)

    # oracle acquisition on the gradient rows (pins the oracle's predict / EI / LCB at this size)
    st = O.GPState(kind, X, Y, variance, ls, noise)
    fo, dfo = st.acquisition("EI", Xg.copy(), with_gradients=True, native=O.ref_native() is not None)
    out["oracle_vs_ref_fmin"] = abs(st.get_fmin() - out["fmin"]) / abs(out["fmin"])
    out["oracle_vs_ref_ei_f"] = rel(fo.ravel(), out["ei_g_f"])
    out["oracle_vs_ref_ei_df"] = rel(dfo, out["ei_g_df"])
    print("  [%s] oracle acquisition: |dfmin|=%.2e |dEI|=%.2e |ddEI|=%.2e" % (name, out["oracle_vs_ref_fmin"],
                                                                             out["oracle_vs_ref_ei_f"], out["oracle_vs_ref_ei_df"]), flush=True)
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)


def main():
    names = sys.argv[1:] or list(CASES)
    ns = rh.load()
    for n in names:
        t0 = time.time()
        run_case(ns, n)
        print("%s done in %.0f s" % (n, time.time() - t0), flush=True)


if __name__ == "__main__":
    main()
