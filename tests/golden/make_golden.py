"""Generate tests/golden/*.npz by executing the reference's OWN numerical source files (see ref_harness.py).

Container-only: needs /root/reference.  Run `python tests/golden/make_golden.py` from the repo root; the .npz files are
committed so that the GPU box (which has no /root/reference) can check the oracle and the CUDA path against them.
Every array below is produced by reference code: GPy kern.K / update_gradients_full / gradients_X, GP.parameters_changed
(ExactGaussianInference.inference), GP.predict / predictive_gradients, GPyOpt GPModel.predict(_withGradients) / get_fmin,
AcquisitionEI / AcquisitionLCB .acquisition_function(_withGradients).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as rh  # noqa: E402

CASES = [
    # name, kind, ard, N, D, M, noise
    ("rbf_ard_n64_d5", "rbf", True, 64, 5, 33, 1e-2),
    ("mat52_ard_n64_d5", "mat52", True, 64, 5, 33, 1e-2),
    ("rbf_iso_n40_d2", "rbf", False, 40, 2, 17, 5e-2),
    ("mat52_iso_n40_d3", "mat52", False, 40, 3, 17, 5e-2),
    ("mat52_ard_n96_d8_exact", "mat52", True, 96, 8, 40, 1e-6),   # the exact_feval noise level (gpmodel.py:72-73)
    ("rbf_ard_n128_d16", "rbf", True, 128, 16, 50, 1e-2),
]


def synth(N, D, M, seed):
    """Same generator family as SURVEY.md 8(d): X~U[0,1], Y = standardised sin(Xw)+0.05 eps."""
    rs = np.random.RandomState(seed)
    X = rs.uniform(0, 1, (N, D))
    w = rs.randn(D)
    Y = np.sin(X @ w)[:, None] + 0.05 * rs.randn(N, 1)
    Y = (Y - Y.mean()) / Y.std()
    Xs = rs.uniform(0, 1, (M, D))
    G_sq = rs.randn(N, N)
    G_rect = rs.randn(M, N)
    return X, Y, Xs, G_sq, G_rect


def run_case(ns, name, kind, ard, N, D, M, noise, seed):
    X, Y, Xs, G_sq, G_rect = synth(N, D, M, seed)
    variance = 1.3
    ls = (0.5 + 0.5 * np.arange(D) / D) if ard else np.array([0.8])
    m = rh.make_model(ns, kind, X, Y, variance, ls, noise, ard=ard)
    k = m.kern
    out = dict(X=X, Y=Y, Xs=Xs, G_sq=G_sq, G_rect=G_rect, variance=variance, lengthscale=ls, noise=noise,
               kind=kind, ard=ard)
    # (a) kernel
    out["K"] = np.array(k.K(X))
    out["K_cross"] = np.array(k.K(Xs, X))
    k.update_gradients_full(G_sq, X)
    out["ugf_sq_var"], out["ugf_sq_len"] = np.array(k.variance.gradient), np.array(k.lengthscale.gradient)
    k.update_gradients_full(G_rect, Xs, X)
    out["ugf_rect_var"], out["ugf_rect_len"] = np.array(k.variance.gradient), np.array(k.lengthscale.gradient)
    out["gX_sq"] = np.array(k.gradients_X(G_sq, X))
    out["gX_rect"] = np.array(k.gradients_X(G_rect, Xs, X))
    # (b) inference (GP.parameters_changed ran in make_model; run again so the kernel gradients are the model's)
    m.parameters_changed()
    out["logL"] = float(m.log_likelihood())
    out["L"] = np.array(m.posterior.woodbury_chol)
    out["alpha"] = np.array(m.posterior.woodbury_vector)
    out["Wi"] = np.array(m.posterior.woodbury_inv)
    out["dL_dK"] = np.array(m.grad_dict["dL_dK"])
    out["grad_var"] = np.array(k.variance.gradient)
    out["grad_len"] = np.array(k.lengthscale.gradient)
    out["grad_noise"] = np.array(m.likelihood.variance.gradient)
    # (c) predict
    mu, var = m.predict(Xs)
    out["pred_mu"], out["pred_var"] = np.array(mu), np.array(var)
    mu, var = m.predict(Xs, include_likelihood=False)
    out["pred_var_noiseless"] = np.array(var)
    mu, cov = m.predict(Xs, full_cov=True)
    out["pred_cov"] = np.array(cov)
    dm, dv = m.predictive_gradients(Xs)
    out["dmu_dX"], out["dv_dX"] = np.array(dm), np.array(dv)
    gm = rh.make_gpmodel(ns, m)
    mm, ss = gm.predict(Xs)
    out["gpm_m"], out["gpm_s"] = np.array(mm), np.array(ss)
    mm, ss, dmdx, dsdx = gm.predict_withGradients(Xs)
    out["gpm_dmdx"], out["gpm_dsdx"] = np.array(dmdx), np.array(dsdx)
    out["fmin"] = float(gm.get_fmin())
    sp = rh._Space()
    ei = ns.AcquisitionEI(gm, sp, optimizer=None, jitter=0.01)
    lcb = ns.AcquisitionLCB(gm, sp, optimizer=None, exploration_weight=2)
    out["ei"] = np.array(ei.acquisition_function(Xs))
    f, df = ei.acquisition_function_withGradients(Xs)
    out["ei_g_f"], out["ei_g_df"] = np.array(f), np.array(df)
    out["lcb"] = np.array(lcb.acquisition_function(Xs))
    f, df = lcb.acquisition_function_withGradients(Xs)
    out["lcb_g_f"], out["lcb_g_df"] = np.array(f), np.array(df)
    # (f) local penalisation (GPyOpt/acquisitions/LP.py): batch of 3 points, L and Min as LocalPenalization would pass them
    Xb, L_lip, Min = Xs[:3].copy(), 2.5, float(Y.min())
    out["lp_Xb"], out["lp_L"], out["lp_Min"] = Xb, L_lip, Min
    for tag, base in (("ei", ei), ("lcb", lcb)):
        lp = ns.AcquisitionLP(gm, sp, None, base)       # LCB switches itself to the softplus transform (LP.py:31-32)
        lp.update_batches(Xb, L_lip, Min)
        out["lp_%s_r" % tag], out["lp_%s_s" % tag] = np.array(lp.r_x0), np.array(lp.s_x0)
        out["lp_%s_f" % tag] = np.array(lp.acquisition_function(Xs[3:]))
        # the reference's gradient broadcasting only works for one point at a time (LP.py:129-132)
        out["lp_%s_df" % tag] = np.vstack([lp.acquisition_function_withGradients(Xs[i:i + 1])[1] for i in range(3, Xs.shape[0])])
        lp.update_batches(None, None, None)
        out["lp_%s_f_nobatch" % tag] = np.array(lp.acquisition_function(Xs[3:]))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    return out


def main():
    ns = rh.load()
    for i, c in enumerate(CASES):
        o = run_case(ns, *c, seed=100 + i)
        print("%-28s logL=%.12g fmin=%.6g" % (c[0], o["logL"], o["fmin"]))


if __name__ == "__main__":
    main()
