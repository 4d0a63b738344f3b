"""GPy.kern surface for the stationary kernels on the hot path: RBF and Matern52 (ARD or isotropic).

Mirrors GPy/GPy/kern/src/kern.py:119-202 (Kern contract), stationary.py:23-370 (Stationary), rbf.py:12-124 (RBF),
stationary.py:546-610 (Matern52).  All arithmetic happens in libgpb200.so (gpb_kern_K, gpb_kern_update_gradients_full,
gpb_kern_gradients_X); there is no NumPy fallback.
"""
import numpy as np

from . import native
from .parameterization import Logexp, Param, Parameterized


class Kern(Parameterized):
    """Base class (kern.py:13-117): input_dim, active_dims (only the identity slicing is supported here)."""

    def __init__(self, input_dim, active_dims, name, useGPU=False, *a, **kw):
        super(Kern, self).__init__(name=name)
        self.input_dim = int(input_dim)
        if active_dims is None:
            active_dims = np.arange(input_dim, dtype=np.int_)
        self.active_dims = np.atleast_1d(np.asarray(active_dims, np.int_))
        assert self.active_dims.size == self.input_dim, \
            "input_dim={} does not match len(active_dim)={}".format(self.input_dim, self.active_dims.size)
        assert np.array_equal(self.active_dims, np.arange(self.input_dim)), \
            "only active_dims = all input dimensions is supported on the B200 path"
        self._all_dims_active = self.active_dims
        self.useGPU = True  # there is no other path

    @property
    def _effective_input_dim(self):
        return self.input_dim

    def K(self, X, X2=None):
        raise NotImplementedError

    def Kdiag(self, X):
        raise NotImplementedError

    def update_gradients_full(self, dL_dK, X, X2=None):
        raise NotImplementedError

    def gradients_X(self, dL_dK, X, X2=None):
        raise NotImplementedError

    def _check(self, X, X2=None):
        # kernel_slice_operations.py:49-51: shape assertions of the slicing wrapper
        assert X.ndim == 2 and X.shape[1] == self.input_dim, \
            "At least {} dimensional X needed, X.shape={!s}".format(self.input_dim, X.shape)
        if X2 is not None:
            assert X2.ndim == 2 and X2.shape[1] == self.input_dim, \
                "At least {} dimensional X2 needed, X2.shape={!s}".format(self.input_dim, X2.shape)


class Stationary(Kern):
    """stationary.py:23-370.  k(r) with r = sqrt(sum_q (x_q - x'_q)^2 / l_q^2)."""

    _kind = None  # "rbf" | "mat52"

    def __init__(self, input_dim, variance, lengthscale, ARD, active_dims, name, useGPU=False, Gower=False, space=None):
        super(Stationary, self).__init__(input_dim, active_dims, name, useGPU=useGPU)
        # the reference's local mixed-variable patch (stationary.py:61-65,116-135): active when Gower and a design space are given
        self.Gower = bool(Gower)
        self.space = space
        self.ARD = bool(ARD)
        if not ARD:
            if lengthscale is None:
                lengthscale = np.ones(1)
            else:
                lengthscale = np.asarray(lengthscale, dtype=np.float64)
                assert lengthscale.size == 1, "Only 1 lengthscale needed for non-ARD kernel"
        else:
            if lengthscale is not None:
                lengthscale = np.asarray(lengthscale, dtype=np.float64)
                assert lengthscale.size in [1, input_dim], "Bad number of lengthscales"
                if lengthscale.size != input_dim:
                    lengthscale = np.ones(input_dim) * lengthscale
            else:
                lengthscale = np.ones(self.input_dim)
        self.lengthscale = Param('lengthscale', lengthscale, Logexp())
        self.variance = Param('variance', variance, Logexp())
        assert self.variance.size == 1
        self.link_parameters(self.variance, self.lengthscale)   # link order = order in m[:] (stationary.py:83)

    def gower_config(self):
        """(continuous dims, discrete dims, ranges of the continuous dims) when the Gower branch of K is active
        (stationary.py:116-119: `self.Gower and self.space is not None`), else None."""
        if not (self.Gower and self.space is not None):
            return None
        return (list(self.space.get_continuous_dims()), list(self.space.get_discrete_dims()), list(self.space.lengthscales()))

    # -- Kern contract ---------------------------------------------------------------------------------------------------
    def K(self, X, X2=None):
        """stationary.py:107-140 (Gower branch :116-135)."""
        X = np.asarray(X, dtype=np.float64)
        X2 = None if X2 is None else np.asarray(X2, dtype=np.float64)
        self._check(X, X2)
        gw = self.gower_config()
        if gw is not None:
            return native.kern_K_gower(self._kind, X, X2, float(self.variance.values[0]), gw)
        return native.kern_K(self._kind, X, X2, float(self.variance.values[0]), self.lengthscale.values)

    def Kdiag(self, X):
        """stationary.py:195-198."""
        ret = np.empty(X.shape[0])
        ret[:] = self.variance.values[0]
        return ret

    def update_gradients_full(self, dL_dK, X, X2=None, reset=True):
        """stationary.py:218-238: fills variance.gradient and lengthscale.gradient."""
        X = np.asarray(X, dtype=np.float64)
        X2 = None if X2 is None else np.asarray(X2, dtype=np.float64)
        self._check(X, X2)
        gw = self.gower_config()
        if gw is not None:   # the patched K only enters the variance term; lengthscale terms stay Euclidean (stationary.py:224-238)
            dv, dl = native.kern_update_gradients_full_gower(self._kind, np.asarray(dL_dK, dtype=np.float64), X, X2,
                                                             float(self.variance.values[0]), self.lengthscale.values, gw)
        else:
            dv, dl = native.kern_update_gradients_full(self._kind, np.asarray(dL_dK, dtype=np.float64), X, X2,
                                                       float(self.variance.values[0]), self.lengthscale.values)
        self.variance.gradient = dv
        self.lengthscale.gradient = dl if self.ARD else dl[0]

    def update_gradients_diag(self, dL_dKdiag, X):
        """stationary.py:206-216."""
        self.variance.gradient = np.sum(dL_dKdiag)
        self.lengthscale.gradient = 0.

    def gradients_X(self, dL_dK, X, X2=None):
        """stationary.py:271-278,354-364."""
        X = np.asarray(X, dtype=np.float64)
        X2 = None if X2 is None else np.asarray(X2, dtype=np.float64)
        self._check(X, X2)
        ls = self.lengthscale.values if self.ARD else np.full(self.input_dim, self.lengthscale.values[0])
        return native.kern_gradients_X(self._kind, np.asarray(dL_dK, dtype=np.float64), X, X2,
                                       float(self.variance.values[0]), ls)

    def gradients_X_diag(self, dL_dKdiag, X):
        """stationary.py:366-367."""
        return np.zeros(X.shape)

    def reset_gradients(self):
        self.variance.gradient = 0.
        self.lengthscale.gradient = 0. if not self.ARD else np.zeros(self.input_dim)

    def input_sensitivity(self, summarize=True):
        return self.variance.values * np.ones(self.input_dim) / self.lengthscale.values ** 2

    def copy(self):
        kw = {"Gower": self.Gower, "space": self.space} if self.Gower else {}
        c = self.__class__(self.input_dim, variance=float(self.variance.values[0]), lengthscale=self.lengthscale.values.copy(),
                           ARD=self.ARD, name=self.name, **kw)
        for src, dst in ((self.variance, c.variance), (self.lengthscale, c.lengthscale)):
            dst._constraint, dst._fixed = src._constraint, src._fixed
        return c

    def __str__(self):
        return "{}(variance={}, lengthscale={}, ARD={})".format(self.name, self.variance.values, self.lengthscale.values, self.ARD)


class RBF(Stationary):
    """k(r) = sigma^2 exp(-r^2 / 2)   (rbf.py:12-54)."""
    _kind = "rbf"

    def __init__(self, input_dim, variance=1., lengthscale=None, ARD=False, active_dims=None, name='rbf', useGPU=False,
                 inv_l=False):
        if inv_l:
            raise NotImplementedError("inv_l parameterisation is outside the B200 hot path")
        super(RBF, self).__init__(input_dim, variance, lengthscale, ARD, active_dims, name, useGPU=useGPU)


class Matern52(Stationary):
    """k(r) = sigma^2 (1 + sqrt(5) r + 5/3 r^2) exp(-sqrt(5) r)   (stationary.py:546-579)."""
    _kind = "mat52"

    def __init__(self, input_dim, variance=1., lengthscale=None, ARD=False, active_dims=None, name='Mat52', useGPU=False,
                 Gower=False, space=None):
        super(Matern52, self).__init__(input_dim, variance, lengthscale, ARD, active_dims, name, useGPU, Gower, space)
