"""gaussian_process_optimization_b200 -- B200-native (sm_100a) exact-GP inner loop behind the GPy / GPyOpt API surface.

The numerical path is hand-written CUDA in libgpb200.so (C ABI: include/gpb200.h); this package is the host-side mirror
of the reference's plug-in interfaces.  There is no CPU fallback.

Drop-in use: replace `import GPy, GPyOpt` by

    from gaussian_process_optimization_b200 import GPy, GPyOpt

and keep `GPy.kern.RBF / Matern52`, `GPy.models.GPRegression`, `GPyOpt.methods.BayesianOptimization`,
`GPyOpt.models.GPModel`, `GPyOpt.acquisitions.AcquisitionEI / AcquisitionLCB`, `GPyOpt.optimization.AcquisitionOptimizer`.
"""
from types import SimpleNamespace as _NS

from . import _lib, native  # noqa: F401
from . import gpyopt as _g
from . import kern as _k
from . import linalg as _la
from . import models as _m
from . import parameterization as _p
from . import sharded as _s

__version__ = "0.1.0"


def set_int8_engine(min_n=8192, digits=8):
    """Experimental, off by default: route the products of the factorisation with at least `min_n` rows (and the predictive products of
    candidate blocks of >= 1024 rows) through the int8 tensor-core engine (csrc/gpb_ozaki.cu).  min_n = 0 switches it off.  digits: 8
    (indistinguishable from the fp64 engine) or 7 (25% faster, agrees to ~1e-13); 10..18 selects the engine's modular mode with that
    many moduli (16: one int8 product per modulus instead of 28 digit-pair products, 56 bits per operand at N = 16384 -- the fastest
    setting; the predictive products then use 18 moduli).  Results stay within the tolerances of tests/, but are not bitwise those of
    the fp64 engine."""
    native.set_ozaki(min_n, digits)

GPy = _NS(
    kern=_NS(Kern=_k.Kern, Stationary=_k.Stationary, RBF=_k.RBF, Matern52=_k.Matern52),
    models=_NS(GPRegression=_m.GPRegression),
    core=_NS(GP=_m.GP, Model=_p.Model, Param=_p.Param, Parameterized=_p.Parameterized,
             parameterization=_NS(Param=_p.Param, Parameterized=_p.Parameterized, transformations=_NS(Logexp=_p.Logexp, Logistic=_p.Logistic))),
    likelihoods=_NS(Gaussian=_m.Gaussian),
    inference=_NS(latent_function_inference=_NS(
        ExactGaussianInference=_m.ExactGaussianInference,
        exact_gaussian_inference=_NS(ExactGaussianInference=_m.ExactGaussianInference),
        posterior=_NS(PosteriorExact=_m.PosteriorExact))),
    util=_NS(linalg=_la),
)

GPyOpt = _NS(
    methods=_NS(BayesianOptimization=_g.BayesianOptimization, ModularBayesianOptimization=_g.ModularBayesianOptimization),
    models=_NS(GPModel=_g.GPModel, BOModel=_g.BOModel, base=_NS(BOModel=_g.BOModel), gpmodel=_NS(GPModel=_g.GPModel)),
    acquisitions=_NS(AcquisitionBase=_g.AcquisitionBase, AcquisitionEI=_g.AcquisitionEI, AcquisitionLCB=_g.AcquisitionLCB,
                     AcquisitionLP=_g.AcquisitionLP),
    optimization=_NS(AcquisitionOptimizer=_g.AcquisitionOptimizer, OptLbfgs=_g.OptLbfgs, apply_optimizer=_g.apply_optimizer,
                     ObjectiveAnchorPointsGenerator=_g.ObjectiveAnchorPointsGenerator,
                     ShardedAnchorScorer=_s.ShardedAnchorScorer),
    core=_NS(BO=_g.BO, evaluators=_NS(Sequential=_g.Sequential, LocalPenalization=_g.LocalPenalization,
                                      batch_local_penalization=_NS(LocalPenalization=_g.LocalPenalization, estimate_L=_g.estimate_L)),
             task=_NS(space=_NS(Design_space=_g.Design_space, bounds_to_space=_g.bounds_to_space),
                      objective=_NS(SingleObjective=_g.SingleObjective), SingleObjective=_g.SingleObjective),
             errors=_NS(InvalidConfigError=_g.InvalidConfigError)),
    experiment_design=_NS(initial_design=_g.initial_design, RandomDesign=_g.RandomDesign),
    util=_NS(general=_NS(normalize=_g.normalize, get_quantiles=_g.get_quantiles, best_value=_g.best_value,
                         samples_multidimensional_uniform=_g.samples_multidimensional_uniform)),
    Design_space=_g.Design_space,
)
