"""GPy.likelihoods.Gaussian, ExactGaussianInference / PosteriorExact, GPy.core.GP and GPy.models.GPRegression, backed by a
model resident on the GPU (gpb_model_* in include/gpb200.h).

Reference interfaces mirrored (paths relative to /root/reference):
  GPy/GPy/likelihoods/gaussian.py:22-110          Gaussian
  GPy/GPy/inference/latent_function_inference/exact_gaussian_inference.py:11-74   ExactGaussianInference.inference
  GPy/GPy/inference/latent_function_inference/posterior.py:8-218,273-302          Posterior / PosteriorExact
  GPy/GPy/core/gp.py:18-110,202-354,407-454,643-664                               GP
  GPy/GPy/models/gp_regression.py:9-36                                            GPRegression
"""
import numpy as np

from . import _lib, native
from .kern import Kern, RBF, Stationary
from .parameterization import Logexp, Model, Param, Parameterized

log_2_pi = np.log(2 * np.pi)


class Gaussian(Parameterized):
    """Gaussian likelihood (gaussian.py:22-110): one parameter, the noise variance."""

    def __init__(self, gp_link=None, variance=1., name='Gaussian_noise'):
        super(Gaussian, self).__init__(name=name)
        self.variance = Param('variance', variance, Logexp())
        self.link_parameter(self.variance)
        self.log_concave = True

    def gaussian_variance(self, Y_metadata=None):
        return self.variance

    def update_gradients(self, grad):
        self.variance.gradient = grad

    def exact_inference_gradients(self, dL_dKdiag, Y_metadata=None):
        return dL_dKdiag.sum()

    def predictive_values(self, mu, var, full_cov=False, Y_metadata=None):
        if full_cov:
            if var.ndim == 2:
                var += np.eye(var.shape[0]) * self.variance.values
            if var.ndim == 3:
                var += np.atleast_3d(np.eye(var.shape[0]) * self.variance.values)
        else:
            var += self.variance.values
        return mu, var

    # paramz exposes constrain_* on Parameterized too: they apply to every parameter below (gpmodel.py:72-76 calls them on
    # `model.Gaussian_noise`)
    def constrain_fixed(self, value=None, warning=True):
        self.variance.constrain_fixed(value, warning)
    fix = constrain_fixed

    def constrain_bounded(self, lower, upper, warning=True):
        self.variance.constrain_bounded(lower, upper, warning)

    def constrain_positive(self, warning=True):
        self.variance.constrain_positive(warning)

    def copy(self):
        c = Gaussian(variance=float(self.variance.values[0]), name=self.name)
        c.variance._constraint, c.variance._fixed = self.variance._constraint, self.variance._fixed
        return c


class StalePosteriorError(_lib.GpbError):
    """A PosteriorExact (or its grad_dict) was used after the resident model it is a view of moved on.

    Deviation from the reference, documented: GPy's Posterior is an immutable snapshot (posterior.py:19-77) holding its own N x N
    arrays; here the N x N state lives once, on the GPU, and a posterior is a VIEW of it.  Every access checks a generation stamp:
    after a later parameters_changed / set_XY -- or a FAILED fit (LinAlgError), after which the reference would still serve its
    last valid posterior -- the old object raises this error instead of returning the new model's numbers (or a mix of both).
    Host copies fetched while the posterior was current (woodbury_chol, woodbury_vector, ...) stay valid and are still served."""


class _LazyGradDict(dict):
    """grad_dict of ExactGaussianInference.inference: 'dL_dK' (N x N) is only materialised when somebody reads it."""

    def __init__(self, nat, dL_dthetaL, alpha_getter):
        super(_LazyGradDict, self).__init__()
        self._nat = nat
        self._gen = nat.generation
        dict.__setitem__(self, 'dL_dthetaL', dL_dthetaL)
        self._alpha_getter = alpha_getter

    def __missing__(self, key):
        if key in ('dL_dK', 'dL_dm') and self._nat.generation != self._gen:
            raise StalePosteriorError("grad_dict['%s'] requested after the model was refitted or its fit failed" % key)
        if key == 'dL_dK':
            v = self._nat.get("dL_dK")
        elif key == 'dL_dm':
            v = self._alpha_getter()
        else:
            raise KeyError(key)
        dict.__setitem__(self, key, v)
        return v

    def keys(self):
        return ['dL_dK', 'dL_dthetaL', 'dL_dm']


class PosteriorExact(object):
    """posterior.py:8-218,273-302 on top of the resident factorisation: woodbury_chol = L, woodbury_vector = alpha,
    woodbury_inv = Ky^-1 are fetched from the device on first access."""

    def __init__(self, nat, kern, X):
        self._native = nat
        self._gen = nat.generation       # stamp of the fit this posterior belongs to
        self._kern = kern
        self._X = X
        self._cache = {}

    @property
    def _nat(self):
        """The resident model, after checking that it still holds THIS posterior (see StalePosteriorError)."""
        if self._native.generation != self._gen:
            raise StalePosteriorError("this posterior is stale: the model behind it was refitted, received new data, or a later fit "
                                      "failed (generation %d, posterior %d); use the model's current .posterior" %
                                      (self._native.generation, self._gen))
        return self._native

    def _get(self, what):
        if what not in self._cache:                       # host copies made while current stay valid (a snapshot, like the reference's)
            self._cache[what] = self._nat.get(what)
        return self._cache[what]

    @property
    def woodbury_chol(self):
        return self._get("L")

    @property
    def woodbury_vector(self):
        return self._get("alpha")

    @property
    def woodbury_inv(self):
        return self._get("Wi")

    @property
    def K(self):
        return self._get("K")
    _K = K

    @property
    def mean(self):
        """posterior.py:79-90: K alpha == the posterior mean at the training inputs."""
        return self._nat.predict(self._X, include_likelihood=False, want_var=False)[0]

    @property
    def covariance(self):
        """posterior.py:92-107: K - K Wi K (the noise-free predictive covariance at the training inputs)."""
        return self._nat.predict_full_cov(self._X, include_likelihood=False)[1]

    def _raw_predict(self, kern, Xnew, pred_var, full_cov=False):
        """posterior.py:273-302 (no clipping here)."""
        Xnew = np.asarray(Xnew, dtype=np.float64)
        if full_cov:
            return self._nat.predict_full_cov(Xnew, include_likelihood=False)
        return self._nat.predict(Xnew, include_likelihood=False)

    # -- what GP / GPModel need beyond the reference's Posterior attributes: the fused device entry points -------------
    def predictive_gradients(self, Xnew, want_var=True):
        """core/gp.py:407-454 -> (dmu_dX (M, D, P), dv_dX (M, D) or None)."""
        return self._nat.predictive_gradients(np.asarray(Xnew, dtype=np.float64), want_var=want_var)

    def set_penalizers(self, transform, Xb, r, s):
        """AcquisitionLP.update_batches (LP.py:40-62): upload the batch points and hammer-function parameters."""
        self._nat.set_penalizers(transform, Xb, r, s)

    def acquisition_lp(self, acq, par, fmin, X, with_gradients=False):
        """AcquisitionLP.acquisition_function(_withGradients) (LP.py:70-140) in one device pass."""
        return self._nat.acquisition_lp(acq, par, fmin, X, with_gradients=with_gradients)

    def fmin(self):
        """min of the posterior mean over the training inputs (gpmodel.py:125-129)."""
        return self._nat.fmin()

    def acquisition(self, acq, par, fmin, X, with_gradients=False, want_moments=False):
        """GPModel.predict(_withGradients) + get_quantiles + EI / LCB (+ gradients) + AcquisitionBase sign, one device pass."""
        return self._nat.acquisition(acq, par, fmin, X, with_gradients=with_gradients, want_moments=want_moments)

    def acq_topk(self, acq, par, fmin, X, k, index_offset=0):
        return self._nat.acq_topk(acq, par, fmin, X, k, index_offset=index_offset)


class ExactGaussianInference(object):
    """exact_gaussian_inference.py:11-74.  `inference` returns (PosteriorExact, log_marginal, grad_dict)."""

    INCREMENTAL = True      # default of `incremental` for instances created without the argument (GPRegression, GPModel)

    def __init__(self, incremental=None):
        self._nat = None
        self._sig = None
        # incremental: when set_XY only appended rows and no hyper-parameter changed, the resident factorisation is extended
        # in O(N^2 b) (gpb_model_append) instead of rebuilt in O(N^3) -- what GPModel.updateModel pays in the reference on
        # every step (gpmodel.py:78-93 -> core/gp.py:202-238,258-271).  Same L, alpha, log-likelihood and gradients up to rounding.
        self.incremental = self.INCREMENTAL if incremental is None else bool(incremental)
        self._resident = None      # (nat, theta, gower key, X) of the last successful jitter-free fit
        self.n_appends = 0

    def on_optimization_start(self):
        pass

    def on_optimization_end(self):
        pass

    def _native_for(self, kern, N, D, P):
        assert isinstance(kern, Stationary), "ExactGaussianInference on the B200 path needs an RBF / Matern52 kernel"
        sig = (kern._kind, kern.ARD, D, P)
        if self._nat is None or self._sig != sig or self._nat.n_cap < N:
            if self._nat is not None:
                self._nat.close()
            cap = max(256, int(1.5 * N))
            cap = (cap + 127) // 128 * 128
            cb = 1024 if cap <= 4096 else 2048
            self._nat = native.NativeModel(kern._kind, kern.ARD, D, P, n_cap=cap, cand_block=cb)
            self._sig = sig
            self._data_id = None
        return self._nat

    def inference(self, kern, X, likelihood, Y, mean_function=None, Y_metadata=None, K=None, variance=None, Z_tilde=None):
        if mean_function is not None or K is not None or Z_tilde is not None:
            raise NotImplementedError("mean functions / precomputed K / EP corrections are outside the B200 hot path")
        X = np.asarray(X, dtype=np.float64)
        Y = np.asarray(Y, dtype=np.float64)
        if variance is None:
            variance = likelihood.gaussian_variance(Y_metadata)
        noise = float(np.asarray(variance).ravel()[0])
        nat = self._native_for(kern, X.shape[0], X.shape[1], Y.shape[1])
        data_id = (id(X), id(Y), X.shape, Y.shape)
        gw = kern.gower_config()
        gw_key = None if gw is None else tuple(map(tuple, gw))
        theta = (float(kern.variance.values[0]), tuple(float(v) for v in kern.lengthscale.values), noise)
        new_data = getattr(self, "_data_id", None) != data_id or getattr(self, "_data_dirty", True)
        info = None
        res = self._resident
        if (new_data and self.incremental and res is not None and res[0] is nat and res[1] == theta and res[2] == gw_key
                and X.shape[0] > res[3].shape[0] and X.shape[0] <= nat.n_cap and nat.n == res[3].shape[0]
                and np.array_equal(X[:res[3].shape[0]], res[3])):
            info, logL, grads = nat.append(X[res[3].shape[0]:], Y, True)
            if info == 0:
                self._data_id, self._data_dirty = data_id, False
                self._data_refs = (X, Y)
                self.n_appends += 1
        if info != 0:
            if new_data:
                nat.set_data(X, Y)
                self._data_id, self._data_dirty = data_id, False
                self._data_refs = (X, Y)   # keep the arrays alive so that id() stays unique
            if getattr(self, "_gower_key", "unset") != (id(nat), gw_key):
                nat.set_gower(gw)
                self._gower_key = (id(nat), gw_key)
            nat.set_theta(theta[0], kern.lengthscale.values, noise)
            # jitchol (linalg.py:56-75): try as is; on failure raise if the diagonal is not positive, else the jitter ladder
            info, logL, grads = nat.fit(True)
        self._resident = (nat, theta, gw_key, X) if info == 0 else None
        if info != 0:
            # mean of diag(Ky) (linalg.py:66): sigma_f^2 for the stationary kernels, sigma_f^(2 d) under the Gower patch, where
            # K(x, x) is a product of d one-dimensional kernels that each carry the variance (stationary.py:116-135)
            kdiag = float(kern.variance.values[0]) ** (X.shape[1] if gw is not None else 1)
            diag = kdiag + noise + 1e-8
            if not diag > 0.:
                raise np.linalg.LinAlgError("not pd: non-positive diagonal elements")
            jitter = diag * 1e-6
            num_tries = 1
            while num_tries <= 5 and np.isfinite(jitter):
                info, logL, grads = nat.fit(True, extra_jitter=jitter)
                if info == 0:
                    break
                jitter *= 10
                num_tries += 1
            if info != 0:
                raise np.linalg.LinAlgError("not positive definite, even with jitter.")
        post = PosteriorExact(nat, kern, X)
        self._last_grads = grads
        gd = _LazyGradDict(nat, grads[-1], lambda: post.woodbury_vector)
        return post, logL, gd

    def invalidate_data(self):
        self._data_dirty = True


class GP(Model):
    """core/gp.py:18-110.  Gaussian likelihood + stationary kernel + exact inference; no mean function, no normaliser."""

    def __init__(self, X, Y, kernel, likelihood, mean_function=None, inference_method=None, name='gp', Y_metadata=None,
                 normalizer=False):
        super(GP, self).__init__(name)
        object.__setattr__(self, "_in_init_", True)
        assert X.ndim == 2
        self.X = np.array(X, dtype=np.float64)
        self.num_data, self.input_dim = self.X.shape
        assert Y.ndim == 2
        if normalizer not in (False, None):
            raise NotImplementedError("output normalisers are outside the B200 hot path (GPyOpt normalises Y itself)")
        self.normalizer = None
        self.Y = np.array(Y, dtype=np.float64)
        self.Y_normalized = self.Y
        assert Y.shape[0] == self.num_data
        _, self.output_dim = self.Y.shape
        assert ((Y_metadata is None) or isinstance(Y_metadata, dict))
        self.Y_metadata = Y_metadata
        assert isinstance(kernel, Kern)
        self.kern = kernel
        assert isinstance(likelihood, Gaussian)
        self.likelihood = likelihood
        if mean_function is not None:
            raise NotImplementedError("mean functions are outside the B200 hot path")
        self.mean_function = None
        if inference_method is None:
            inference_method = ExactGaussianInference()
        self.inference_method = inference_method
        self.link_parameter(self.kern)
        self.link_parameter(self.likelihood)
        self.posterior = None
        self._log_marginal_likelihood = None
        self.grad_dict = None
        object.__setattr__(self, "_in_init_", False)
        self._trigger()          # paramz runs parameters_changed once the hierarchy is connected

    # -- data ------------------------------------------------------------------------------------------------------------
    def set_XY(self, X=None, Y=None):
        """core/gp.py:202-238."""
        self._updates = False
        if Y is not None:
            self.Y = np.array(Y, dtype=np.float64)
            self.Y_normalized = self.Y
            _, self.output_dim = self.Y.shape
        if X is not None:
            self.X = np.array(X, dtype=np.float64)
            self.num_data, self.input_dim = self.X.shape
        if hasattr(self.inference_method, "invalidate_data"):
            self.inference_method.invalidate_data()
        self.update_model(True)

    def set_X(self, X):
        self.set_XY(X=X)

    def set_Y(self, Y):
        self.set_XY(Y=Y)

    # -- inference -------------------------------------------------------------------------------------------------------
    def parameters_changed(self):
        """core/gp.py:258-271."""
        self.posterior, self._log_marginal_likelihood, self.grad_dict = self.inference_method.inference(
            self.kern, self.X, self.likelihood, self.Y_normalized, self.mean_function, self.Y_metadata)
        g = getattr(self.inference_method, "_last_grads", None)
        if g is not None and isinstance(self.kern, Stationary):
            # the fused device pass already produced every gradient (never materialising dL_dK)
            self.likelihood.update_gradients(g[-1])
            self.kern.variance.gradient = g[0]
            self.kern.lengthscale.gradient = g[1:-1].copy() if self.kern.ARD else g[1]
        else:
            self.likelihood.update_gradients(self.grad_dict['dL_dthetaL'])
            self.kern.update_gradients_full(self.grad_dict['dL_dK'], self.X)

    def log_likelihood(self):
        return self._log_marginal_likelihood

    @property
    def _predictive_variable(self):
        return self.X

    # -- prediction ------------------------------------------------------------------------------------------------------
    def _raw_predict(self, Xnew, full_cov=False, kern=None):
        """core/gp.py:279-295."""
        return self.posterior._raw_predict(kern=self.kern if kern is None else kern, Xnew=Xnew,
                                           pred_var=self._predictive_variable, full_cov=full_cov)

    def predict(self, Xnew, full_cov=False, Y_metadata=None, kern=None, likelihood=None, include_likelihood=True):
        """core/gp.py:297-354."""
        mean, var = self._raw_predict(Xnew, full_cov=full_cov, kern=kern)
        if include_likelihood:
            if likelihood is None:
                likelihood = self.likelihood
            mean, var = likelihood.predictive_values(mean, var, full_cov, Y_metadata=Y_metadata)
        return mean, var

    def predict_noiseless(self, Xnew, full_cov=False, Y_metadata=None, kern=None):
        return self.predict(Xnew, full_cov, Y_metadata, kern, None, False)

    def predictive_gradients(self, Xnew, kern=None, want_var=True):
        """core/gp.py:407-454 -> (dmu_dX (M, D, P), dv_dX (M, D)).  want_var=False (not in the reference) skips the variance
        gradient for callers that discard it (estimate_L, batch_local_penalization.py:56-58) and returns None in its place."""
        if want_var:
            return self.posterior.predictive_gradients(np.asarray(Xnew, dtype=np.float64))
        return self.posterior.predictive_gradients(np.asarray(Xnew, dtype=np.float64), want_var=False)

    def posterior_covariance_between_points(self, X1, X2):
        """posterior.py:109-130 through one joint full-covariance prediction."""
        X1, X2 = np.asarray(X1, dtype=np.float64), np.asarray(X2, dtype=np.float64)
        _, cov = self.posterior._raw_predict(self.kern, np.vstack([X1, X2]), self._predictive_variable, full_cov=True)
        return cov[:X1.shape[0], X1.shape[0]:]

    def copy(self):
        return GP(self.X.copy(), self.Y.copy(), self.kern.copy(), self.likelihood.copy(), name=self.name,
                  inference_method=type(self.inference_method)())

    def __str__(self):
        return "{}: N={}, D={}, log-likelihood={}\n  {}\n  noise variance={}".format(
            self.name, self.num_data, self.input_dim, self._log_marginal_likelihood, self.kern, self.likelihood.variance.values)


class GPRegression(GP):
    """models/gp_regression.py:9-36."""

    def __init__(self, X, Y, kernel=None, Y_metadata=None, normalizer=None, noise_var=1., mean_function=None):
        if kernel is None:
            kernel = RBF(X.shape[1])
        likelihood = Gaussian(variance=noise_var)
        super(GPRegression, self).__init__(X, Y, kernel, likelihood, name='GP regression', Y_metadata=Y_metadata,
                                           normalizer=normalizer, mean_function=mean_function)

    def copy(self):
        m = GPRegression(self.X.copy(), self.Y.copy(), kernel=self.kern.copy(), noise_var=float(self.likelihood.variance.values[0]))
        src, dst = self.likelihood.variance, m.likelihood.variance
        dst._constraint, dst._fixed = src._constraint, src._fixed
        return m
