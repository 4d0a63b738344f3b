"""Test double: the CPU oracle behind the SAME host-side interfaces the CUDA path sits behind.

`OracleInference` is a drop-in for `models.ExactGaussianInference` (GP accepts `inference_method=`, core/gp.py:38,99-105) and
`OraclePosterior` for `models.PosteriorExact`; `OracleGPModel` is `gpyopt.GPModel` building its GP on them.  Running the
identical host logic (parameter transforms, L-BFGS-B, anchor generation, BO loop, NumPy RNG order) once on the CUDA numerics
and once on the oracle numerics is how the tests compare optimiser runs and BO trajectories.

Test infrastructure only -- never imported by the product package.
"""
import numpy as np

from gaussian_process_optimization_b200 import gpyopt, kern as _kern, models
from oracle import gp_oracle as O


class OraclePosterior(object):
    def __init__(self, kind, ard, X, variance, lengthscale, noise, post, gower=None):
        self.kind, self.ard, self.X = kind, ard, X
        self.variance, self.lengthscale, self.noise, self.post, self.gower = variance, lengthscale, noise, post, gower

    woodbury_chol = property(lambda self: self.post.woodbury_chol)
    woodbury_vector = property(lambda self: self.post.woodbury_vector)
    woodbury_inv = property(lambda self: self.post.woodbury_inv)
    K = property(lambda self: self.post.K)

    def _a(self):
        return self.kind, self.post, self.X

    def _raw_predict(self, kern, Xnew, pred_var, full_cov=False):
        return O.raw_predict(self.kind, self.post, self.X, np.asarray(Xnew, dtype=np.float64), self.variance, self.lengthscale,
                             self.ard, full_cov, self.gower)

    def predictive_gradients(self, Xnew, want_var=True):
        return O.predictive_gradients(self.kind, self.post, self.X, Xnew, self.variance, self.lengthscale, self.ard, gower=self.gower)

    def fmin(self):
        return O.gpmodel_get_fmin(self.kind, self.post, self.X, self.variance, self.lengthscale, self.noise, self.ard, gower=self.gower)

    def acquisition(self, acq, par, fmin, X, with_gradients=False, want_moments=False):
        X = np.atleast_2d(np.asarray(X, dtype=np.float64))
        r = {}
        if with_gradients:
            m, s, dmdx, dsdx = O.gpmodel_predict_withGradients(self.kind, self.post, self.X, X, self.variance, self.lengthscale,
                                                               self.noise, self.ard, gower=self.gower)
            if want_moments:
                r["m"], r["s"], r["dmdx"], r["dsdx"] = m, s.copy(), dmdx, dsdx
            f, df = (O.acq_EI(m, s, fmin, par, dmdx, dsdx) if acq == "EI" else O.acq_LCB(m, s, par, dmdx, dsdx))
            r["f"], r["df"] = -f, -df
        else:
            m, s = O.gpmodel_predict(self.kind, self.post, self.X, X, self.variance, self.lengthscale, self.noise, self.ard,
                                     gower=self.gower)
            if want_moments:
                r["m"], r["s"] = m, s.copy()
            f = O.acq_EI(m, s, fmin, par) if acq == "EI" else O.acq_LCB(m, s, par)
            r["f"] = -f
        return r

    def acq_topk(self, acq, par, fmin, X, k, index_offset=0):
        f = self.acquisition(acq, par, fmin, X)["f"].ravel()
        order = np.argsort(f, kind="stable")[:k]
        return f[order], order.astype(np.int64) + index_offset, np.asarray(X)[order]


class OracleInference(object):
    """exact_gaussian_inference.py:37-74 on the CPU oracle (the jitter ladder is inside O.pdinv -> O.jitchol)."""

    def on_optimization_start(self):
        pass

    def on_optimization_end(self):
        pass

    def inference(self, kern, X, likelihood, Y, mean_function=None, Y_metadata=None, K=None, variance=None, Z_tilde=None):
        v = float(kern.variance.values[0])
        ls = kern.lengthscale.values.copy()
        noise = float(np.asarray(likelihood.gaussian_variance()).ravel()[0])
        gw = kern.gower_config()
        logL, grads, post = O.log_likelihood_and_gradients(kern._kind, X, Y, v, ls, noise, ard=kern.ARD, gower=gw)
        self._last_grads = grads
        gd = {"dL_dthetaL": grads[-1], "dL_dm": post.woodbury_vector}
        return OraclePosterior(kern._kind, kern.ARD, X, v, ls, noise, post, gw), logL, gd


def oracle_gp_regression(X, Y, kernel, noise_var=1.):
    return models.GP(X, Y, kernel, models.Gaussian(variance=noise_var), inference_method=OracleInference(), name="GP regression")


class OracleGPModel(gpyopt.GPModel):
    """GPyOpt GPModel whose GP runs on the oracle (same class otherwise: updateModel, optimize_restarts, predict, ...)."""

    batched_rows_bitwise = False    # NumPy / BLAS pick different kernels for 1-row and 5-row products: keep the sequential anchor loop

    def _create_model(self, X, Y):
        self.input_dim = X.shape[1]
        if self.kernel is None:
            kern = _kern.Matern52(self.input_dim, variance=1., ARD=self.ARD, Gower=self.Gower, space=self.space)
        else:
            kern = self.kernel
            self.kernel = None
        noise_var = Y.var() * 0.01 if self.noise_var is None else self.noise_var
        self.model = oracle_gp_regression(X, Y, kern, noise_var)
        if self.exact_feval:
            self.model.Gaussian_noise.constrain_fixed(1e-6, warning=False)
        else:
            self.model.Gaussian_noise.constrain_bounded(1e-9, 1e6, warning=False)
