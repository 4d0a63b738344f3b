"""CPU ORACLE -- test infrastructure, NOT product code.

A NumPy/SciPy restatement of the exact-GP inner loop that GPy 1.9.6 / GPyOpt 1.2.5 (vendored under /root/reference) run on
every Bayesian-optimisation step.  Every function cites the reference file:line it follows and uses the same LAPACK/BLAS
entry points (dpotrf, dpotri, dpotrs, dtrtrs, dsyrk) in the same order, so that it reproduces the reference's floating-point
behaviour as closely as a restatement can.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import this module.
The product package (`gaussian_process_optimization_b200`) never does: it fails loudly when its CUDA library is missing.

Pinning (see DESIGN.md "Oracle"):
  * numerics (K, gradients, inference, predict, predictive gradients, EI/LCB):  PINNED against the reference's own source
    files executed in the build container through tests/golden/ref_harness.py; the resulting vectors are committed as
    tests/golden/*.npz (generator: tests/golden/make_golden.py) and re-checked by tests/test_oracle_golden.py.
  * paramz behaviour (Logexp/Logistic transforms, optimize / optimize_restarts / randomize): paramz is an un-vendored,
    unpinned dependency (GPy/setup.py:162 `paramz>=0.9.0`) that is absent from /root/reference -> restated from its
    published algorithm; PARITY UNPINNED for the optimiser trajectory.

Conventions: X (N,D), Y (N,P), theta = (variance sigma_f^2, lengthscale (D,) or (1,), noise sigma_n^2); kind in {"rbf","mat52"}.
"""
import ctypes
import os

import numpy as np
from scipy import linalg
from scipy.linalg import blas, lapack
from scipy.special import erfc

LOG_2_PI = np.log(2 * np.pi)  # exact_gaussian_inference.py:9

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF_SO = os.path.join(_HERE, "_ref", "libstationary_utils_ref.so")
_ref_lib = None


def ref_native():
    """The reference's own stationary_utils.c compiled by oracle/Makefile (None if not built)."""
    global _ref_lib
    if _ref_lib is None and os.path.exists(_REF_SO):
        _ref_lib = ctypes.CDLL(_REF_SO)
    return _ref_lib


# ----------------------------------------------------------------------------------------------------------------------
# L0c linalg  (GPy/GPy/util/linalg.py, diag.py, linalg_cython.pyx)
# ----------------------------------------------------------------------------------------------------------------------
def symmetrify(A, upper=False):
    """linalg_cython.pyx:9-18 / linalg.py:356-379: copy lower->upper (upper=False) or upper->lower, in place."""
    if not upper:
        iu = np.triu_indices_from(A, k=1)
        A[iu] = A.T[iu]
    else:
        il = np.tril_indices_from(A, k=-1)
        A[il] = A.T[il]


def tdot(mat):
    """linalg.py:299-323 tdot_blas: dsyrk(upper of the F-ordered view) + symmetrify(upper=True)."""
    nn = mat.shape[0]
    out = np.zeros((nn, nn))
    matF = np.asfortranarray(mat)
    out = blas.dsyrk(alpha=1.0, a=matF, beta=0.0, c=out, overwrite_c=1, trans=0, lower=0)
    symmetrify(out, upper=True)
    return np.ascontiguousarray(out)


def jitchol(A, maxtries=5):
    """linalg.py:56-75.  dpotrf(lower=1); on failure the jitter ladder diag.mean()*1e-6 * 10^t, t<maxtries."""
    A = np.ascontiguousarray(A)
    L, info = lapack.dpotrf(A, lower=1)
    if info == 0:
        return L
    diagA = np.diag(A)
    if np.any(diagA <= 0.):
        raise linalg.LinAlgError("not pd: non-positive diagonal elements")
    jitter = diagA.mean() * 1e-6
    num_tries = 1
    while num_tries <= maxtries and np.isfinite(jitter):
        try:
            return linalg.cholesky(A + np.eye(A.shape[0]) * jitter, lower=True)
        except Exception:
            jitter *= 10
        finally:
            num_tries += 1
    raise linalg.LinAlgError("not positive definite, even with jitter.")


def force_F_ordered(A):
    """linalg.py:31-38"""
    if A.flags['F_CONTIGUOUS']:
        return A
    return np.asfortranarray(A)


def dtrtrs(A, B, lower=1, trans=0, unitdiag=0):
    """linalg.py:95-114"""
    A = np.asfortranarray(A)
    return lapack.dtrtrs(A, B, lower=lower, trans=trans, unitdiag=unitdiag)


def dpotrs(A, B, lower=1):
    """linalg.py:116-125"""
    A = force_F_ordered(A)
    return lapack.dpotrs(A, B, lower=lower)


def dpotri(A, lower=1):
    """linalg.py:127-145 (dpotri + symmetrify)"""
    A = force_F_ordered(A)
    R, info = lapack.dpotri(A, lower=lower)
    symmetrify(R)
    return R, info


def dtrtri(L):
    """linalg.py:217-227"""
    L = force_F_ordered(L)
    return lapack.dtrtri(L, lower=1)[0]


def pdinv(A, with_Li=True):
    """linalg.py:193-214.  `with_Li=False` skips the dtrtri whose result exact inference never uses."""
    L = jitchol(A)
    logdet = 2. * np.sum(np.log(np.diag(L)))
    Li = dtrtri(L) if with_Li else None
    Ai, _ = dpotri(L, lower=1)
    symmetrify(Ai)
    return Ai, L, Li, logdet


# ----------------------------------------------------------------------------------------------------------------------
# L0b kernels  (GPy/GPy/kern/src/stationary.py, rbf.py)
# ----------------------------------------------------------------------------------------------------------------------
def _ls(lengthscale):
    return np.atleast_1d(np.asarray(lengthscale, dtype=np.float64))


def unscaled_dist(X, X2=None):
    """stationary.py:155-173"""
    if X2 is None:
        Xsq = np.sum(np.square(X), 1)
        r2 = -2. * tdot(X) + (Xsq[:, None] + Xsq[None, :])
        r2[np.diag_indices_from(r2)] = 0.
        r2 = np.clip(r2, 0, np.inf)
        return np.sqrt(r2)
    X1sq = np.sum(np.square(X), 1)
    X2sq = np.sum(np.square(X2), 1)
    r2 = -2. * np.dot(X, X2.T) + (X1sq[:, None] + X2sq[None, :])
    r2 = np.clip(r2, 0, np.inf)
    return np.sqrt(r2)


def scaled_dist(X, X2, lengthscale, ard):
    """stationary.py:176-193"""
    ls = _ls(lengthscale)
    if ard:
        if X2 is not None:
            X2 = X2 / ls
        return unscaled_dist(X / ls, X2)
    return unscaled_dist(X, X2) / ls


def K_of_r(kind, r, variance):
    """rbf.py:50-51 ; stationary.py:575-576"""
    if kind == "rbf":
        return variance * np.exp(-0.5 * r ** 2)
    if kind == "mat52":
        return variance * (1 + np.sqrt(5.) * r + 5. / 3 * r ** 2) * np.exp(-np.sqrt(5.) * r)
    raise ValueError(kind)


def dK_dr(kind, r, variance):
    """rbf.py:53-54 ; stationary.py:578-579"""
    if kind == "rbf":
        return -r * K_of_r(kind, r, variance)
    if kind == "mat52":
        return variance * (10. / 3 * r - 5. * r - 5. * np.sqrt(5.) / 3 * r ** 2) * np.exp(-np.sqrt(5.) * r)
    raise ValueError(kind)


def K_gower(kind, X, X2, variance, gower):
    """stationary.py:116-135, the local "Gower" patch of the reference: a product of one-dimensional kernels, r = |dx| / range
    on the continuous dimensions (space.lengthscales(), GPyOpt core/task/space.py:351-362) and r = [x != x'] on the discrete
    ones.  Every factor carries the variance; the kernel's own lengthscale parameter is ignored.
    gower = (continuous dims, discrete dims, ranges of the continuous dims in that order)."""
    const_dims, disc_dims, ranges = gower
    if X2 is None:
        X2 = X
    numDims = X.shape[1]
    K_1D = [None for _ in range(numDims)]
    for index, const_dim in enumerate(const_dims):
        r = abs(X[:, np.newaxis, const_dim] - X2[np.newaxis, :, const_dim]) / ranges[index]
        K_1D[const_dim] = K_of_r(kind, r, variance)
    for disc_dim in disc_dims:
        r = (X[:, np.newaxis, disc_dim] != X2[np.newaxis, :, disc_dim]).astype(int)
        K_1D[disc_dim] = K_of_r(kind, r, variance)
    kernel = K_1D[0]
    for dim in range(numDims - 1):
        kernel = kernel * K_1D[dim + 1]
    return kernel


def K(kind, X, X2, variance, lengthscale, ard=True, gower=None):
    """stationary.py:107-140 (Gower branch :116-135 when gower is given, else :137-139)"""
    if gower is not None:
        return K_gower(kind, X, X2, variance, gower)
    return K_of_r(kind, scaled_dist(X, X2, lengthscale, ard), variance)


def Kdiag(X, variance):
    """stationary.py:195-198"""
    ret = np.empty(X.shape[0])
    ret[:] = variance
    return ret


def inv_dist(X, X2, lengthscale, ard):
    """stationary.py:251-258"""
    dist = scaled_dist(X, X2, lengthscale, ard).copy()
    with np.errstate(divide="ignore"):
        return 1. / np.where(dist != 0., dist, np.inf)


def lengthscale_grads_pure(tmp, X, X2, lengthscale):
    """stationary.py:260-261"""
    ls = _ls(lengthscale)
    D = X.shape[1]
    return -np.array([np.sum(tmp * np.square(X[:, q:q + 1] - X2[:, q:q + 1].T)) for q in range(D)]) / ls ** 3


def lengthscale_grads_native(tmp, X, X2, lengthscale):
    """stationary.py:263-269 with the reference's compiled C loop (stationary_utils.c:34-48, same loop nest / order as
    stationary_cython.pyx:51-60)."""
    lib = ref_native()
    if lib is None:
        raise RuntimeError("oracle/_ref not built (make -C oracle)")
    ls = _ls(lengthscale)
    N, M = tmp.shape
    Q = X.shape[1]
    X, X2, tmp = np.ascontiguousarray(X), np.ascontiguousarray(X2), np.ascontiguousarray(tmp)
    grads = np.zeros(Q)
    dp = ctypes.POINTER(ctypes.c_double)
    lib._lengthscale_grads(ctypes.c_int(N), ctypes.c_int(M), ctypes.c_int(Q), tmp.ctypes.data_as(dp),
                           X.ctypes.data_as(dp), X2.ctypes.data_as(dp), grads.ctypes.data_as(dp))
    return -grads / ls ** 3


def update_gradients_full(kind, dL_dK, X, X2, variance, lengthscale, ard=True, native=False, gower=None):
    """stationary.py:218-238 -> (d/dvariance, d/dlengthscale).  With the Gower patch only self.K changes (:224); the
    lengthscale part keeps using the Euclidean scaled distance (:227-238), exactly as the reference does."""
    ls = _ls(lengthscale)
    dvar = np.sum(K(kind, X, X2, variance, ls, ard, gower) * dL_dK) / variance
    r = scaled_dist(X, X2, ls, ard)
    dL_dr = dK_dr(kind, r, variance) * dL_dK
    if ard:
        tmp = dL_dr * inv_dist(X, X2, ls, ard)
        if X2 is None:
            X2 = X
        dlen = (lengthscale_grads_native if native else lengthscale_grads_pure)(tmp, X, X2, ls)
    else:
        dlen = -np.sum(dL_dr * r) / ls
    return dvar, np.atleast_1d(dlen)


def gradients_X(kind, dL_dK, X, X2, variance, lengthscale, ard=True, native=False):
    """stationary.py:271-278,336-364 (pure: :336-352; cython: :354-364 -> stationary_utils.c:1-14)."""
    ls = _ls(lengthscale)
    invdist = inv_dist(X, X2, ls, ard)
    dL_dr = dK_dr(kind, scaled_dist(X, X2, ls, ard), variance) * dL_dK
    tmp = invdist * dL_dr
    if X2 is None:
        tmp = tmp + tmp.T
        X2 = X
    if native and ref_native() is not None:
        Xc, X2c, tmpc = np.ascontiguousarray(X), np.ascontiguousarray(X2), np.ascontiguousarray(tmp)
        grad = np.zeros(Xc.shape)
        dp = ctypes.POINTER(ctypes.c_double)
        ref_native()._grad_X(ctypes.c_int(Xc.shape[0]), ctypes.c_int(Xc.shape[1]), ctypes.c_int(X2c.shape[0]),
                             Xc.ctypes.data_as(dp), X2c.ctypes.data_as(dp), tmpc.ctypes.data_as(dp),
                             grad.ctypes.data_as(dp))
    else:
        grad = np.empty(X.shape, dtype=np.float64)
        for q in range(X.shape[1]):
            np.sum(tmp * (X[:, q][:, None] - X2[:, q][None, :]), axis=1, out=grad[:, q])
    return grad / ls ** 2


def gradients_X_diag(X):
    """stationary.py:366-367"""
    return np.zeros(X.shape)


# ----------------------------------------------------------------------------------------------------------------------
# L0a inference  (exact_gaussian_inference.py, posterior.py, likelihoods/gaussian.py)
# ----------------------------------------------------------------------------------------------------------------------
class Posterior(object):
    """What PosteriorExact carries (posterior.py:19-77): woodbury_chol L, woodbury_vector alpha, K; lazy woodbury_inv."""

    def __init__(self, L, alpha, Kmat, Wi=None):
        self.woodbury_chol = L
        self.woodbury_vector = alpha
        self.K = Kmat
        self._woodbury_inv = Wi

    @property
    def woodbury_inv(self):
        """posterior.py:176-196 (dpotri + symmetrify; numerically the same matrix pdinv already produced)."""
        if self._woodbury_inv is None:
            self._woodbury_inv, _ = dpotri(self.woodbury_chol, lower=1)
            symmetrify(self._woodbury_inv)
        return self._woodbury_inv


def exact_inference(kind, X, Y, variance, lengthscale, noise, ard=True, with_Li=False, gower=None):
    """exact_gaussian_inference.py:37-74 -> (Posterior, log_marginal, {'dL_dK','dL_dthetaL','dL_dm'})."""
    Kmat = K(kind, X, None, variance, lengthscale, ard, gower)
    Ky = Kmat.copy()
    Ky[np.diag_indices_from(Ky)] += noise + 1e-8                      # :55-56
    Wi, LW, LWi, W_logdet = pdinv(Ky, with_Li=with_Li)               # :58
    alpha, _ = dpotrs(LW, Y, lower=1)                                 # :60
    log_marginal = 0.5 * (-Y.size * LOG_2_PI - Y.shape[1] * W_logdet - np.sum(alpha * Y))   # :62
    dL_dK = 0.5 * (tdot(alpha) - Y.shape[1] * Wi)                     # :70
    dL_dthetaL = np.diag(dL_dK).sum()                                 # :72 ; gaussian.py:78-79
    return Posterior(LW, alpha, Kmat, Wi), log_marginal, {'dL_dK': dL_dK, 'dL_dthetaL': dL_dthetaL, 'dL_dm': alpha}


def log_likelihood_and_gradients(kind, X, Y, variance, lengthscale, noise, ard=True, native=False, gower=None):
    """GP.parameters_changed (core/gp.py:258-271): inference, then likelihood + kernel gradient updates.

    Returns (logL, grads) with grads ordered like m[:] = [kern.variance, kern.lengthscale..., Gaussian_noise.variance]
    (link order: stationary.py:83, core/gp.py:108-109)."""
    post, logL, gd = exact_inference(kind, X, Y, variance, lengthscale, noise, ard, gower=gower)
    dvar, dlen = update_gradients_full(kind, gd['dL_dK'], X, None, variance, lengthscale, ard, native=native, gower=gower)
    return logL, np.concatenate([[dvar], dlen, [gd['dL_dthetaL']]]), post


def raw_predict(kind, post, X, Xnew, variance, lengthscale, ard=True, full_cov=False, gower=None):
    """posterior.py:273-302 (PosteriorExact._raw_predict); no clipping here."""
    Kx = K(kind, X, Xnew, variance, lengthscale, ard, gower)
    mu = np.dot(Kx.T, post.woodbury_vector)
    if mu.ndim == 1:
        mu = mu.reshape(-1, 1)
    if full_cov:
        Kxx = K(kind, Xnew, None, variance, lengthscale, ard, gower)
        tmp = dtrtrs(post.woodbury_chol, Kx)[0]
        var = Kxx - tdot(tmp.T)
    else:
        Kxx = Kdiag(Xnew, variance)
        tmp = dtrtrs(post.woodbury_chol, Kx)[0]
        var = (Kxx - np.square(tmp).sum(0))[:, None]
    return mu, var


def predict(kind, post, X, Xnew, variance, lengthscale, noise, ard=True, full_cov=False, include_likelihood=True, gower=None):
    """core/gp.py:297-354 (+ gaussian.py:102-110 predictive_values)."""
    mu, var = raw_predict(kind, post, X, Xnew, variance, lengthscale, ard, full_cov, gower)
    if include_likelihood:
        if full_cov:
            var = var + np.eye(var.shape[0]) * noise
        else:
            var = var + noise
    return mu, var


def predictive_gradients(kind, post, X, Xnew, variance, lengthscale, ard=True, native=False, gower=None):
    """core/gp.py:407-454 -> (dmu_dX (M,D,P), dv_dX (M,D))."""
    P = post.woodbury_vector.shape[1]
    mean_jac = np.empty((Xnew.shape[0], Xnew.shape[1], P))
    for i in range(P):
        mean_jac[:, :, i] = gradients_X(kind, post.woodbury_vector[:, i:i + 1].T, Xnew, X, variance, lengthscale, ard,
                                        native=native)
    dv_dX = gradients_X_diag(Xnew)
    alpha = -2. * np.dot(K(kind, Xnew, X, variance, lengthscale, ard, gower), post.woodbury_inv)
    dv_dX = dv_dX + gradients_X(kind, alpha, Xnew, X, variance, lengthscale, ard, native=native)
    return mean_jac, dv_dX


# ----------------------------------------------------------------------------------------------------------------------
# L2/L3  GPyOpt adaptor + acquisitions (models/gpmodel.py, util/general.py, acquisitions/{EI,LCB,base}.py)
# ----------------------------------------------------------------------------------------------------------------------
def gpmodel_predict(kind, post, X, Xnew, variance, lengthscale, noise, ard=True, with_noise=True, gower=None):
    """gpmodel.py:95-112: clip v at 1e-10 AFTER adding the noise, return (m, sqrt(v))."""
    if Xnew.ndim == 1:
        Xnew = Xnew[None, :]
    m, v = predict(kind, post, X, Xnew, variance, lengthscale, noise, ard, include_likelihood=with_noise, gower=gower)
    v = np.clip(v, 1e-10, np.inf)
    return m, np.sqrt(v)


def gpmodel_get_fmin(kind, post, X, variance, lengthscale, noise, ard=True, gower=None):
    """gpmodel.py:125-129"""
    return predict(kind, post, X, X, variance, lengthscale, noise, ard, gower=gower)[0].min()


def gpmodel_predict_withGradients(kind, post, X, Xnew, variance, lengthscale, noise, ard=True, native=False, gower=None):
    """gpmodel.py:131-142"""
    if Xnew.ndim == 1:
        Xnew = Xnew[None, :]
    m, v = predict(kind, post, X, Xnew, variance, lengthscale, noise, ard, gower=gower)
    v = np.clip(v, 1e-10, np.inf)
    dmdx, dvdx = predictive_gradients(kind, post, X, Xnew, variance, lengthscale, ard, native=native, gower=gower)
    dmdx = dmdx[:, :, 0]
    dsdx = dvdx / (2 * np.sqrt(v))
    return m, np.sqrt(v), dmdx, dsdx


def get_quantiles(acquisition_par, fmin, m, s):
    """util/general.py:113-128 (mutates s in place, like the reference)."""
    if isinstance(s, np.ndarray):
        s[s < 1e-10] = 1e-10
    elif s < 1e-10:
        s = 1e-10
    u = (fmin - m - acquisition_par) / s
    phi = np.exp(-0.5 * u ** 2) / np.sqrt(2 * np.pi)
    Phi = 0.5 * erfc(-u / np.sqrt(2))
    return phi, Phi, u


def acq_EI(m, s, fmin, jitter=0.01, dmdx=None, dsdx=None):
    """acquisitions/EI.py:32-51"""
    phi, Phi, u = get_quantiles(jitter, fmin, m, s)
    f_acqu = s * (u * Phi + phi)
    if dmdx is None:
        return f_acqu
    return f_acqu, dsdx * phi - Phi * dmdx


def acq_LCB(m, s, exploration_weight=2, dmdx=None, dsdx=None):
    """acquisitions/LCB.py:31-46"""
    f_acqu = -m + exploration_weight * s
    if dmdx is None:
        return f_acqu
    return f_acqu, -dmdx + exploration_weight * dsdx


class GPState(object):
    """Frozen GP (kind, data, theta, posterior): what GPModel holds after updateModel.  Convenience for tests/bench."""

    def __init__(self, kind, X, Y, variance, lengthscale, noise, ard=True, gower=None):
        self.kind, self.X, self.Y, self.ard = kind, np.ascontiguousarray(X, dtype=np.float64), np.asarray(Y, float), ard
        self.variance, self.lengthscale, self.noise = float(variance), _ls(lengthscale).copy(), float(noise)
        self.gower = gower
        self.post, self.logL, self.grad_dict = exact_inference(kind, self.X, self.Y, self.variance, self.lengthscale,
                                                               self.noise, ard, gower=gower)
        self._fmin = None

    def _a(self):
        return (self.kind, self.post, self.X)

    def predict(self, Xnew, with_noise=True):
        return gpmodel_predict(self.kind, self.post, self.X, Xnew, self.variance, self.lengthscale, self.noise, self.ard,
                               with_noise, gower=self.gower)

    def predict_withGradients(self, Xnew, native=False):
        return gpmodel_predict_withGradients(self.kind, self.post, self.X, Xnew, self.variance, self.lengthscale,
                                             self.noise, self.ard, native=native, gower=self.gower)

    def get_fmin(self):
        # the reference recomputes this on every acquisition call (gpmodel.py:125-129); the value only depends on the model
        if self._fmin is None:
            self._fmin = gpmodel_get_fmin(self.kind, self.post, self.X, self.variance, self.lengthscale, self.noise,
                                          self.ard, gower=self.gower)
        return self._fmin

    def acquisition(self, acq, Xnew, par=None, with_gradients=False, native=False):
        """AcquisitionBase.acquisition_function(_withGradients) (acquisitions/base.py:33-50) with the constant cost
        (core/task/cost.py:76-80) and an unconstrained space (indicator == 1): returns -acq [, -dacq]."""
        if Xnew.ndim == 1:
            Xnew = Xnew[None, :]
        if not with_gradients:
            m, s = self.predict(Xnew)
            if acq == "EI":
                f = acq_EI(m, s, self.get_fmin(), 0.01 if par is None else par)
            else:
                f = acq_LCB(m, s, 2 if par is None else par)
            return -f
        if acq == "EI":
            fmin = self.get_fmin()
            m, s, dmdx, dsdx = self.predict_withGradients(Xnew, native=native)
            f, df = acq_EI(m, s, fmin, 0.01 if par is None else par, dmdx, dsdx)
        else:
            m, s, dmdx, dsdx = self.predict_withGradients(Xnew, native=native)
            f, df = acq_LCB(m, s, 2 if par is None else par, dmdx, dsdx)
        return -f, -df


# ----------------------------------------------------------------------------------------------------------------------
# paramz restatement (UNPINNED -- source not under /root/reference)
# ----------------------------------------------------------------------------------------------------------------------
_LIM_VAL = 36.0
_LOG_LIM_VAL = np.log(np.finfo(np.float64).max)


class Logexp(object):
    """paramz.transformations.Logexp: theta = log(1+exp(x))."""

    @staticmethod
    def f(x):
        x = np.asarray(x, dtype=np.float64)
        return np.where(x > _LIM_VAL, x, np.log1p(np.exp(np.clip(x, -_LOG_LIM_VAL, _LIM_VAL))))

    @staticmethod
    def finv(f):
        f = np.asarray(f, dtype=np.float64)
        return np.where(f > _LIM_VAL, f, np.log(np.expm1(f)))

    @staticmethod
    def gradfactor(f, df):
        f = np.asarray(f, dtype=np.float64)
        return df * np.where(f > _LIM_VAL, 1., -np.expm1(-f))


class Logistic(object):
    """paramz.transformations.Logistic(lower, upper) (from constrain_bounded)."""

    def __init__(self, lower, upper):
        self.lower, self.upper = float(lower), float(upper)
        self.difference = self.upper - self.lower

    def f(self, x):
        x = np.array(x, dtype=np.float64)
        x[x < -300.] = -300.
        return self.lower + self.difference / (1. + np.exp(-x))

    def finv(self, f):
        f = np.asarray(f, dtype=np.float64)
        return np.log(np.clip(f - self.lower, 1e-10, np.inf) / np.clip(self.upper - f, 1e-10, np.inf))

    def gradfactor(self, f, df):
        f = np.asarray(f, dtype=np.float64)
        return df * (f - self.lower) * (self.upper - f) / self.difference
