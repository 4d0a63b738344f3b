"""Run the UNMODIFIED reference numerics (GPy 1.9.6 / GPyOpt 1.2.5 under /root/reference) in this container.

Test infrastructure only (used by tests/golden/make_golden.py and by the container-only cross-check tests).
`/root/reference` does not exist on the GPU box, so nothing on the `-m gpu` path imports this module.

Why a harness: `import GPy` cannot work here -- `paramz` is an un-vendored, absent dependency and
GPy/__init__.py trips over NumPy-2 / Python-3.12 removals.  The *numerical* modules themselves are fine, so we
  * register skeleton packages (`GPy`, `GPy.util`, `GPy.kern.src`, ...) whose __path__ points into /root/reference but whose
    __init__.py is never executed,
  * provide tiny stand-ins for the framework glue only (paramz Param/Parameterized/Cache_this/ObsAr, the config flag, the
    psi-statistics helpers, the two Cython extension modules),
  * and then import the reference's own source files for everything that does arithmetic:
      GPy/util/linalg.py, GPy/util/diag.py, GPy/kern/src/{kern,kernel_slice_operations,stationary,rbf}.py,
      GPy/inference/latent_function_inference/{exact_gaussian_inference,posterior}.py, GPy/likelihoods/gaussian.py,
      GPy/core/gp.py, GPy/models/gp_regression.py, GPyOpt/util/general.py, GPyOpt/acquisitions/{base,EI,LCB,LP}.py,
      GPyOpt/models/gpmodel.py.
The Cython stand-ins call the reference's own C file (GPy/GPy/kern/src/stationary_utils.c) compiled into
oracle/_ref/libstationary_utils_ref.so by oracle/Makefile, or restate the 10-line .pyx loops in NumPy.
"""
import ctypes
import warnings
import importlib
import os
import sys
import types

import numpy as np

REF = os.environ.get("GPB200_REFERENCE", "/root/reference")
_REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
_REF_SO = os.path.join(_REPO, "oracle", "_ref", "libstationary_utils_ref.so")


def available():
    return os.path.isdir(os.path.join(REF, "GPy", "GPy")) and os.path.isdir(os.path.join(REF, "GPyOpt", "GPyOpt"))


# ----------------------------------------------------------------------------------------------------------------------
# framework-glue stand-ins (NOT numerics)
# ----------------------------------------------------------------------------------------------------------------------
class Param(np.ndarray):
    """Minimal paramz.Param: an ndarray with a name and a .gradient slot."""

    def __new__(cls, name, value, transform=None, *a, **kw):
        obj = np.atleast_1d(np.array(value, dtype=np.float64)).view(cls)
        obj.name = name
        obj.gradient = np.zeros(obj.shape)
        obj._transform = transform
        return obj

    def __array_finalize__(self, obj):
        self.name = getattr(obj, "name", None)
        self.gradient = getattr(obj, "gradient", None)
        self._transform = getattr(obj, "_transform", None)

    @property
    def values(self):
        return self.view(np.ndarray)

    def __array_wrap__(self, out, context=None, return_scalar=False):
        # arithmetic on a Param yields a plain ndarray, like paramz
        out = np.asarray(out)
        return out[()] if return_scalar else out

    def fix(self, value=None, warning=True):
        if value is not None:
            self[...] = value
    constrain_fixed = fix

    def constrain_bounded(self, lo, hi, warning=True):
        pass


class Parameterized(object):
    def __init__(self, name=None, *a, **kw):
        self.name = name
        self.parameters = []

    def link_parameter(self, p, index=None):
        self.parameters.append(p)

    def link_parameters(self, *ps):
        for p in ps:
            self.link_parameter(p)

    def unlink_parameter(self, p):
        self.parameters = [q for q in self.parameters if q is not p]

    def update_model(self, *a):
        return True

    def parameters_changed(self):
        pass


class Logexp(object):
    pass


def Cache_this(*a, **kw):
    def deco(f):
        return f
    return deco


class ObsAr(np.ndarray):
    """paramz.ObsAr stand-in: a float64 copy of the data."""

    def __new__(cls, x):
        return np.array(x, dtype=np.float64).view(cls)

    def __array_wrap__(self, out, context=None, return_scalar=False):
        out = np.asarray(out)
        return out[()] if return_scalar else out

    def copy(self):
        return ObsAr(np.asarray(self))


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _pkg(name, path):
    m = types.ModuleType(name)
    m.__path__ = [path]
    m.__package__ = name
    sys.modules[name] = m
    parent, _, child = name.rpartition(".")
    if parent:
        setattr(sys.modules[parent], child, m)
    return m


_loaded = None


class _IdList(list):
    """paramz keeps parameters in an ArrayList whose `in` tests identity (GP.set_XY asks `self.X in self.parameters`)."""

    def __contains__(self, item):
        return any(item is e for e in self)


def _full_paramz():
    """The repo's own restatement of paramz (gaussian_process_optimization_b200/parameterization.py: transforms, observer
    pattern, optimize / optimize_restarts / randomize) as the stand-in, so that the reference's GPModel.updateModel and BO
    loop can run end to end.  Host logic only -- it does no GP arithmetic."""
    sys.path.insert(0, _REPO)
    from gaussian_process_optimization_b200 import parameterization as P

    class FullParameterized(P.Parameterized):
        def __init__(self, name=None, *a, **kw):
            super(FullParameterized, self).__init__(name, *a, **kw)
            self.parameters = _IdList()

        def unlink_parameter(self, param):
            self.parameters = _IdList(p for p in self.parameters if p is not param)
            param._parent = None

        # paramz Parameterized forwards constrain_* to every parameter below it (gpmodel.py:72-76 calls them on Gaussian_noise)
        def constrain_fixed(self, value=None, warning=True):
            for p in self.flattened_parameters():
                p.constrain_fixed(value, warning)
        fix = constrain_fixed

        def constrain_bounded(self, lower, upper, warning=True):
            for p in self.flattened_parameters():
                p.constrain_bounded(lower, upper, warning)

        def constrain_positive(self, warning=True):
            for p in self.flattened_parameters():
                p.constrain_positive(warning)

    class FullModel(P.Model, FullParameterized):
        def __init__(self, name):
            P.Model.__init__(self, name)
            self.parameters = _IdList()

    return P.Param, FullParameterized, FullModel, P.Logexp


def load(full_paramz=False):
    """Install the stand-ins, import the reference's numerical modules, return a namespace of them.
    full_paramz=True swaps the minimal Param/Parameterized stand-ins for the repo's paramz restatement (needed to run
    optimize_restarts and the BO loop; tests/golden/ref_bo_harness.py)."""
    global _loaded, Param, Parameterized, Logexp
    if _loaded is not None:
        return _loaded
    ModelBase = None
    if full_paramz:
        Param, Parameterized, ModelBase, Logexp = _full_paramz()
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF)
    warnings.filterwarnings("ignore", category=SyntaxWarning)
    if not os.path.exists(_REF_SO):
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(_REPO, "oracle")])

    # NumPy-2 / SciPy removals the 2018 sources rely on (aliases only)
    if not hasattr(np, "bool"):
        np.bool = bool
    if not hasattr(np.linalg, "linalg"):
        np.linalg.linalg = np.linalg

    # --- paramz stand-in -------------------------------------------------------------------------------------------
    pz = _mod("paramz", ObsAr=ObsAr, Param=Param, Parameterized=Parameterized)
    pz.caching = _mod("paramz.caching", Cache_this=Cache_this)
    pz.transformations = _mod("paramz.transformations", Logexp=Logexp, __fixed__="fixed")
    pz.parameterized = _mod("paramz.parameterized", ParametersChangedMeta=type, Parameterized=Parameterized)

    # --- GPy skeleton ----------------------------------------------------------------------------------------------
    g = os.path.join(REF, "GPy", "GPy")
    _pkg("GPy", g)
    _pkg("GPy.util", os.path.join(g, "util"))
    _pkg("GPy.kern", os.path.join(g, "kern"))
    _pkg("GPy.kern.src", os.path.join(g, "kern", "src"))
    core = _pkg("GPy.core", os.path.join(g, "core"))
    par = _pkg("GPy.core.parameterization", os.path.join(g, "core", "parameterization"))
    _pkg("GPy.inference", os.path.join(g, "inference"))
    lfi = _pkg("GPy.inference.latent_function_inference", os.path.join(g, "inference", "latent_function_inference"))
    lik = _pkg("GPy.likelihoods", os.path.join(g, "likelihoods"))
    models = _pkg("GPy.models", os.path.join(g, "models"))

    class _Cfg(object):
        def getboolean(self, section, key):
            return True  # [cython] working = True, GPy/GPy/defaults.cfg:26-27 -> native code paths are taken
    util_config = _mod("GPy.util.config", config=_Cfg())
    sys.modules["GPy.util"].config = util_config

    par.Param = Param
    par.Parameterized = Parameterized
    par.parameterized = _mod("GPy.core.parameterization.parameterized", Parameterized=Parameterized)

    class VariationalPosterior(object):
        pass
    par.variational = _mod("GPy.core.parameterization.variational", VariationalPosterior=VariationalPosterior)
    core.Param = Param
    core.Parameterized = Parameterized

    # Cython extension stand-ins ------------------------------------------------------------------------------------
    lib = ctypes.CDLL(_REF_SO)
    dp = ctypes.POINTER(ctypes.c_double)

    def _p(a):
        assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
        return a.ctypes.data_as(dp)

    def grad_X(N, D, M, X, X2, tmp, grad):
        # stationary_cython.pyx:19-29 -> stationary_utils.c:1-14 (the reference's own C, compiled as is)
        tmp = np.ascontiguousarray(tmp)
        lib._grad_X(ctypes.c_int(N), ctypes.c_int(D), ctypes.c_int(M), _p(X), _p(X2), _p(tmp), _p(grad))

    def lengthscale_grads(N, M, Q, tmp, X, X2, grad):
        # stationary_cython.pyx:51-60 is a plain serial triple loop; stationary_utils.c:34-48 (_lengthscale_grads) is the
        # same loop nest with the same summation order (n outer, m inner, per q) -> use the compiled reference C.
        tmp = np.ascontiguousarray(tmp)
        lib._lengthscale_grads(ctypes.c_int(N), ctypes.c_int(M), ctypes.c_int(Q), _p(tmp), _p(X), _p(X2), _p(grad))

    sc = _mod("GPy.kern.src.stationary_cython", grad_X=grad_X, lengthscale_grads=lengthscale_grads)
    sys.modules["GPy.kern.src"].stationary_cython = sc

    def symmetrify(A, upper):
        # linalg_cython.pyx:9-18
        if not upper:
            iu = np.triu_indices_from(A, k=1)
            A[iu] = A.T[iu]
        else:
            il = np.tril_indices_from(A, k=-1)
            A[il] = A.T[il]
    lc = _mod("GPy.util.linalg_cython", symmetrify=symmetrify)
    sys.modules["GPy.util"].linalg_cython = lc

    class _Dummy(object):
        def __init__(self, *a, **kw):
            pass
    psi = _mod("GPy.kern.src.psi_comp", PSICOMP_RBF=_Dummy, PSICOMP_RBF_GPU=_Dummy, PSICOMP_GH=_Dummy)
    sys.modules["GPy.kern.src"].psi_comp = psi
    gk = _mod("GPy.kern.src.grid_kerns", GridRBF=_Dummy)
    sys.modules["GPy.kern.src"].grid_kerns = gk

    class LatentFunctionInference(object):
        # GPy/inference/latent_function_inference/__init__.py:36-46 (hooks called by GP.optimize, no-ops for exact inference)
        def on_optimization_start(self):
            pass

        def on_optimization_end(self):
            pass
    lfi.LatentFunctionInference = LatentFunctionInference

    # --- the reference's own numerical modules ---------------------------------------------------------------------
    ns = types.SimpleNamespace()
    ns.diag = importlib.import_module("GPy.util.diag")
    ns.linalg = importlib.import_module("GPy.util.linalg")
    ns.kern_base = importlib.import_module("GPy.kern.src.kern")
    ns.stationary = importlib.import_module("GPy.kern.src.stationary")
    ns.rbf = importlib.import_module("GPy.kern.src.rbf")
    ns.posterior = importlib.import_module("GPy.inference.latent_function_inference.posterior")
    ns.egi = importlib.import_module("GPy.inference.latent_function_inference.exact_gaussian_inference")

    class Identity(object):
        pass
    lik.link_functions = _mod("GPy.likelihoods.link_functions", Identity=Identity)

    class Likelihood(Parameterized):
        def __init__(self, gp_link, name):
            Parameterized.__init__(self, name)
            self.gp_link = gp_link
    lik.likelihood = _mod("GPy.likelihoods.likelihood", Likelihood=Likelihood)
    ns.gaussian = importlib.import_module("GPy.likelihoods.gaussian")
    lik.Gaussian = ns.gaussian.Gaussian
    lik.Likelihood = Likelihood

    class MixedNoise(Likelihood):
        pass
    lik.MixedNoise = MixedNoise

    kern_pkg = sys.modules["GPy.kern"]
    kern_pkg.Kern = ns.kern_base.Kern
    kern_pkg.RBF = ns.rbf.RBF
    kern_pkg.Matern52 = ns.stationary.Matern52

    if ModelBase is not None:
        Model = ModelBase
    else:
        class Model(Parameterized):
            pass
    core.model = _mod("GPy.core.model", Model=Model)
    core.Model = Model

    class Mapping(object):
        pass
    core.mapping = _mod("GPy.core.mapping", Mapping=Mapping)

    class EP(object):
        pass
    lfi.expectation_propagation = _mod("GPy.inference.latent_function_inference.expectation_propagation", EP=EP)
    importlib.import_module("GPy.util.normalizer")
    ns.gp = importlib.import_module("GPy.core.gp")
    core.GP = ns.gp.GP
    ns.gp_regression = importlib.import_module("GPy.models.gp_regression")
    models.GPRegression = ns.gp_regression.GPRegression

    # --- GPyOpt skeleton -------------------------------------------------------------------------------------------
    o = os.path.join(REF, "GPyOpt", "GPyOpt")
    _pkg("GPyOpt", o)
    _pkg("GPyOpt.util", os.path.join(o, "util"))
    _pkg("GPyOpt.core", os.path.join(o, "core"))
    _pkg("GPyOpt.core.task", os.path.join(o, "core", "task"))
    _pkg("GPyOpt.models", os.path.join(o, "models"))
    _pkg("GPyOpt.acquisitions", os.path.join(o, "acquisitions"))
    importlib.import_module("GPyOpt.core.errors")
    ns.general = importlib.import_module("GPyOpt.util.general")
    importlib.import_module("GPyOpt.models.base")
    ns.gpmodel = importlib.import_module("GPyOpt.models.gpmodel")
    sys.modules["GPyOpt.models"].GPModel = ns.gpmodel.GPModel
    sys.modules["GPyOpt.models"].GPModel_MCMC = ns.gpmodel.GPModel_MCMC   # isinstance test in core/bo.py:99
    ns.cost = importlib.import_module("GPyOpt.core.task.cost")
    ns.acq_base = importlib.import_module("GPyOpt.acquisitions.base")
    ns.EI = importlib.import_module("GPyOpt.acquisitions.EI")
    ns.LCB = importlib.import_module("GPyOpt.acquisitions.LCB")

    class AcquisitionLCB_MCMC(ns.LCB.AcquisitionLCB):   # only referenced in an isinstance test of LP.py:33
        pass
    _mod("GPyOpt.acquisitions.LCB_mcmc", AcquisitionLCB_MCMC=AcquisitionLCB_MCMC)
    ns.LP = importlib.import_module("GPyOpt.acquisitions.LP")
    ns.AcquisitionLP = ns.LP.AcquisitionLP

    ns.RBF = ns.rbf.RBF
    ns.Matern52 = ns.stationary.Matern52
    ns.GPRegression = ns.gp_regression.GPRegression
    ns.GPModel = ns.gpmodel.GPModel
    ns.AcquisitionEI = ns.EI.AcquisitionEI
    ns.AcquisitionLCB = ns.LCB.AcquisitionLCB
    _loaded = ns
    return ns


class _Space(object):
    """Stand-in for GPyOpt Design_space: unconstrained box (indicator_constraints == 1, core/task/space.py:303-318)."""

    def indicator_constraints(self, x):
        return np.ones((np.atleast_2d(x).shape[0], 1))


def make_model(ns, kind, X, Y, variance, lengthscale, noise, ard=True):
    """Reference GPRegression with the given hyper-parameters; runs the reference's GP.parameters_changed once."""
    D = X.shape[1]
    K = ns.RBF if kind == "rbf" else ns.Matern52
    k = K(D, variance=variance, lengthscale=lengthscale, ARD=ard)
    m = ns.GPRegression(X.copy(), Y.copy(), kernel=k, noise_var=noise)
    m.parameters_changed()  # GPy/GPy/core/gp.py:258-271 (paramz would trigger this on link)
    return m


def make_gpmodel(ns, m):
    gm = ns.GPModel(exact_feval=False, verbose=False)
    gm.model = m
    return gm
