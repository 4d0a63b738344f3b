"""Golden vectors for the reference's local "Gower" mixed-variable kernel patch (GPy/GPy/kern/src/stationary.py:61-65,116-135),
the configuration run.py actually uses (Gower=True, run.py:1207-1224).  Container-only; every array is produced by reference code:
Matern52(Gower=True, space=<reference Design_space>), GPRegression, GPModel, AcquisitionEI / LCB / LP."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as rh  # noqa: E402
import ref_bo_harness as rb  # noqa: E402

CASES = [
    # name, ARD, N, M, noise, variance
    ("gower_mat52_iso_n60", False, 60, 25, 1e-6, 1.0),     # what GPModel builds for run.py: variance 1, exact_feval noise
    ("gower_mat52_ard_n48_var", True, 48, 20, 1e-3, 1.4),  # variance != 1 exposes the variance^D product and the Kdiag mismatch
]
DOMAIN = [{'name': 'a', 'type': 'discrete', 'domain': (0, 1, 2, 3)}, {'name': 'x', 'type': 'continuous', 'domain': (-5., 10.)},
          {'name': 'b', 'type': 'discrete', 'domain': (1, 2, 3)}, {'name': 'y', 'type': 'continuous', 'domain': (1., 15.)}]


def main():
    ns = rb.load_bo()
    space = ns.space.Design_space(DOMAIN, None)
    D = 4
    for ci, (name, ard, N, M, noise, variance) in enumerate(CASES):
        rs = np.random.RandomState(500 + ci)
        def draw(n):
            return np.c_[rs.choice(DOMAIN[0]['domain'], n), rs.uniform(-5, 10, n), rs.choice(DOMAIN[2]['domain'], n), rs.uniform(1, 15, n)].astype(float)
        X, Xs = draw(N), draw(M)
        Xs[:3] = X[:3]                                   # coincident rows: r = 0 on every dimension
        Y = (np.sin(X[:, 1] / 3) + 0.3 * X[:, 0] - 0.2 * (X[:, 2] == 2) + 0.01 * X[:, 3] ** 2)[:, None]
        Y = (Y - Y.mean()) / Y.std()
        ls = (np.array([0.8, 2.0, 1.1, 3.0]) if ard else np.array([1.7]))
        k = ns.Matern52(D, variance=variance, lengthscale=ls, ARD=ard, Gower=True, space=space)
        m = ns.GPRegression(X.copy(), Y.copy(), kernel=k, noise_var=noise)
        m.parameters_changed()
        G_sq, G_rect = rs.randn(N, N), rs.randn(M, N)
        out = dict(X=X, Y=Y, Xs=Xs, G_sq=G_sq, G_rect=G_rect, variance=variance, lengthscale=ls, noise=noise, ard=ard,
                   cont_dims=np.array(space.get_continuous_dims()), disc_dims=np.array(space.get_discrete_dims()),
                   ranges=np.array(space.lengthscales(), dtype=float))
        out["K"], out["K_cross"] = np.array(k.K(X)), np.array(k.K(Xs, X))
        k.update_gradients_full(G_sq, X)
        out["ugf_sq_var"], out["ugf_sq_len"] = np.array(k.variance.gradient), np.array(k.lengthscale.gradient)
        k.update_gradients_full(G_rect, Xs, X)
        out["ugf_rect_var"], out["ugf_rect_len"] = np.array(k.variance.gradient), np.array(k.lengthscale.gradient)
        m.parameters_changed()
        out["logL"] = float(m.log_likelihood())
        out["L"], out["alpha"] = np.array(m.posterior.woodbury_chol), np.array(m.posterior.woodbury_vector)
        out["grad_var"], out["grad_len"] = np.array(k.variance.gradient), np.array(k.lengthscale.gradient)
        out["grad_noise"] = np.array(m.likelihood.variance.gradient)
        mu, var = m.predict(Xs)
        out["pred_mu"], out["pred_var"] = np.array(mu), np.array(var)
        dm, dv = m.predictive_gradients(Xs)
        out["dmu_dX"], out["dv_dX"] = np.array(dm), np.array(dv)
        gm = rh.make_gpmodel(ns, m)
        mm, ss, dmdx, dsdx = gm.predict_withGradients(Xs)
        out["gpm_m"], out["gpm_s"], out["gpm_dmdx"], out["gpm_dsdx"] = np.array(mm), np.array(ss), np.array(dmdx), np.array(dsdx)
        out["fmin"] = float(gm.get_fmin())
        sp = rh._Space()
        ei = ns.AcquisitionEI(gm, sp, optimizer=None, jitter=0.01)
        lcb = ns.AcquisitionLCB(gm, sp, optimizer=None, exploration_weight=2)
        f, df = ei.acquisition_function_withGradients(Xs)
        out["ei_f"], out["ei_df"] = np.array(f), np.array(df)
        f, df = lcb.acquisition_function_withGradients(Xs)
        out["lcb_f"], out["lcb_df"] = np.array(f), np.array(df)
        lp = ns.AcquisitionLP(gm, sp, None, ei)
        Xb = Xs[5:8].copy()
        lp.update_batches(Xb, 1.5, float(Y.min()))
        out["lp_Xb"], out["lp_f"] = Xb, np.array(lp.acquisition_function(Xs[8:]))
        np.savez_compressed(os.path.join(HERE, "gower", name + ".npz"), **out)
        print("%-28s logL=%.12g fmin=%.6g  K[0,0]=%.6g Kdiag=%.6g" % (name, out["logL"], out["fmin"], out["K"][0, 0], variance))


if __name__ == "__main__":
    main()
