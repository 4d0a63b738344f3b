"""GPyOpt surface for the Bayesian-optimisation step on top of the B200 GP model.

Mirrors (paths relative to /root/reference/GPyOpt/GPyOpt):
  models/base.py:7-33, models/gpmodel.py:9-177            BOModel, GPModel
  acquisitions/base.py:6-68, EI.py:7-51, LCB.py:7-46      AcquisitionBase, AcquisitionEI, AcquisitionLCB
  util/general.py:113-128,203-234                         get_quantiles, normalize
  optimization/optimizer.py:28-61,130-168                 OptLbfgs, apply_optimizer
  optimization/anchor_points_generator.py:8-98            ObjectiveAnchorPointsGenerator
  optimization/acquisition_optimizer.py:16-77             AcquisitionOptimizer
  experiment_design/random_design.py:7-91                 RandomDesign (np.random consumption order preserved)
  core/task/space.py, variables.py                        Design_space with continuous / discrete variables
  core/task/objective.py:24-76                            SingleObjective
  core/evaluators/sequential.py:7-23                      Sequential
  acquisitions/LP.py:10-140                               AcquisitionLP (local penalisation)
  core/evaluators/batch_local_penalization.py:9-70        LocalPenalization, estimate_L
  core/bo.py:20-260, methods/bayesian_optimization.py:76-202   BO, BayesianOptimization
Host logic only (bookkeeping, RNG order, L-BFGS-B through SciPy like the reference); every GP quantity comes from the device.

ATTRIBUTION.  This file is a host-side mirror of GPyOpt 1.2.5's control plane, kept only so that the BO-trajectory parity tests
(BASELINE config 1) can run on a GPU box where /root/reference does not exist.  Class / method names, signatures, attribute names
and -- for the bookkeeping classes (DuplicateManager, RandomDesign, Design_space, normalize, BO.run_optimization,
BayesianOptimization.__init__ and its choosers) -- the statement order of the bodies follow GPyOpt closely, because the NumPy RNG
consumption order and the stopping rules are part of the trajectory being compared.  GPyOpt is
    Copyright (c) 2015, the GPyOpt authors.  All rights reserved.  Licensed under the BSD 3-clause license
    (GPyOpt/LICENSE.txt in the reference tree: redistribution in source and binary forms, with or without modification, is
    permitted provided that the copyright notice, the list of conditions and the disclaimer are retained; the names of the authors
    may not be used to endorse derived products; the software is provided "as is", without warranty of any kind).
It is NOT part of the B200 hot path and earns no credit as such (SURVEY.md section 2 marks the control plane out of scope); the product
is the C ABI underneath (include/gpb200.h).  What is original here: LockstepEvaluator / the lock-step anchor refinement, the
sharded / distributed variants, and the device-backed GPModel / acquisition plumbing.
"""
import time

import numpy as np
from scipy.special import erfc

from . import kern as _kern
from .models import GPRegression


# ----------------------------------------------------------------------------------------------------------------------
# errors / general utilities
# ----------------------------------------------------------------------------------------------------------------------
class InvalidConfigError(Exception):
    pass


class FullyExploredOptimizationDomainError(Exception):
    pass


def normalize(Y, normalization_type='stats'):
    """util/general.py:203-234."""
    Y = np.asarray(Y, dtype=float)
    if np.max(Y.shape) != Y.size:
        raise NotImplementedError('Only 1-dimensional arrays are supported.')
    if normalization_type == 'stats':
        Y_norm = Y - Y.mean()
        std = Y.std()
        if std > 0:
            Y_norm /= std
    elif normalization_type == 'maxmin':
        Y_norm = Y - Y.min()
        y_range = np.ptp(Y)
        if y_range > 0:
            Y_norm /= y_range
            Y_norm = 2 * (Y_norm - 0.5)
    else:
        raise ValueError('Unknown normalization type: {}'.format(normalization_type))
    return Y_norm


def best_value(Y, sign=1):
    """util/general.py:131-143."""
    n = Y.shape[0]
    Y_best = np.ones(n)
    for i in range(n):
        Y_best[i] = Y[:(i + 1)].min() if sign == 1 else Y[:(i + 1)].max()
    return Y_best


def get_quantiles(acquisition_par, fmin, m, s):
    """util/general.py:113-128 -- only used for user-supplied (non-B200) models; the GPModel path fuses this on the device."""
    if isinstance(s, np.ndarray):
        s[s < 1e-10] = 1e-10
    elif s < 1e-10:
        s = 1e-10
    u = (fmin - m - acquisition_par) / s
    phi = np.exp(-0.5 * u ** 2) / np.sqrt(2 * np.pi)
    Phi = 0.5 * erfc(-u / np.sqrt(2))
    return (phi, Phi, u)


class DuplicateManager(object):
    """util/duplicate_manager.py:7-41: the set of configurations already evaluated / pending / black-listed."""

    def __init__(self, space, zipped_X, pending_zipped_X=None, ignored_zipped_X=None):
        self.space = space
        self.unique_points = set()
        self.unique_points.update(tuple(x.flatten()) for x in zipped_X)
        if np.any(pending_zipped_X):
            self.unique_points.update(tuple(x.flatten()) for x in pending_zipped_X)
        if np.any(ignored_zipped_X):
            self.unique_points.update(tuple(x.flatten()) for x in ignored_zipped_X)

    def is_zipped_x_duplicate(self, zipped_x):
        return tuple(zipped_x.flatten()) in self.unique_points

    def is_unzipped_x_duplicate(self, unzipped_x):
        return self.is_zipped_x_duplicate(self.space.zip_inputs(np.atleast_2d(unzipped_x)))


def constant_cost_withGradients(x):
    """core/task/cost.py:76-80."""
    return np.ones(x.shape[0])[:, None], np.zeros(x.shape)


# ----------------------------------------------------------------------------------------------------------------------
# design space
# ----------------------------------------------------------------------------------------------------------------------
class _Variable(object):
    def __init__(self, name, var_type, domain):
        self.name, self.type, self.domain = name, var_type, domain
        self.dimensionality = 1
        self.dimensionality_in_model = 1

    def is_continuous(self):
        return self.type == 'continuous'

    def is_discrete(self):
        return self.type == 'discrete'

    def is_bandit(self):
        return False

    def get_bounds(self):
        if self.type == 'continuous':
            return [tuple(self.domain)]
        return [(min(self.domain), max(self.domain))]

    def round(self, value_array):
        """variables.py:124-139 (continuous: clamp) and :186-199 (discrete: nearest domain value, first wins)."""
        v = value_array[0]
        if self.type == 'continuous':
            lo, hi = self.domain[0], self.domain[1]
            if v < lo:
                v = lo
            elif v > hi:
                v = hi
            return [v]
        r = self.domain[0]
        for dv in self.domain:
            if np.abs(dv - v) < np.abs(r - v):
                r = dv
        return [r]


class Design_space(object):
    """core/task/space.py:14-100 for continuous and discrete variables (categorical / bandit arms are out of scope)."""

    supported_types = ['continuous', 'discrete']

    def __init__(self, space, constraints=None, store_noncontinuous=False):
        self.config_space = space
        self.space_expanded = []
        for i, d in enumerate(space):
            t = d.get('type', 'continuous')
            if t not in self.supported_types:
                raise InvalidConfigError("variable type %r is not supported on the B200 path" % t)
            dim = int(d.get('dimensionality', 1))
            name = d.get('name', 'var_' + str(i + 1))
            for j in range(dim):
                self.space_expanded.append(_Variable(name if dim == 1 else '{}_{}'.format(name, j + 1), t, d['domain']))
        self.space = self.space_expanded
        self.config_space_expanded = [{'name': v.name, 'type': v.type, 'domain': v.domain, 'dimensionality': 1}
                                      for v in self.space_expanded]
        self.dimensionality = len(self.space_expanded)
        self.objective_dimensionality = self.dimensionality
        self.model_dimensionality = self.dimensionality
        self.model_input_dims = [1] * self.dimensionality
        if constraints is not None:
            for c in constraints:
                if 'constrain' in c:
                    c['constraint'] = c['constrain']
        self.constraints = constraints

    def has_constraints(self):
        return self.constraints is not None

    def has_continuous(self):
        return any(v.is_continuous() for v in self.space)

    def has_discrete(self):
        return any(v.is_discrete() for v in self.space)

    def _has_bandit(self):
        return False

    def get_bounds(self):
        b = []
        for v in self.space_expanded:
            b += v.get_bounds()
        return b

    def lengthscales(self):
        """core/task/space.py:351-362 (local patch of the reference): the domain width of every continuous variable, in order --
        the per-dimension scales of the Gower kernel."""
        return [v.domain[-1] - v.domain[0] for v in self.space if v.type == 'continuous']

    def get_continuous_bounds(self):
        return [tuple(v.domain) for v in self.space if v.type == 'continuous']

    def get_continuous_dims(self):
        return [i for i, v in enumerate(self.space_expanded) if v.type == 'continuous']

    def get_discrete_dims(self):
        return [i for i, v in enumerate(self.space_expanded) if v.type == 'discrete']

    def input_dim(self):
        return self.dimensionality

    def unzip_inputs(self, X):
        return np.atleast_2d(X)

    def zip_inputs(self, X):
        return np.atleast_2d(X)

    def indicator_constraints(self, x):
        """core/task/space.py:303-318 (constraints are python expressions in `x`)."""
        x = np.atleast_2d(x)
        I_x = np.ones((x.shape[0], 1))
        if self.constraints is not None:
            for d in self.constraints:
                constraint = eval('lambda x:' + d['constraint'], {'np': np})
                ind_x = (constraint(x) < 0) * 1
                I_x *= ind_x.reshape(x.shape[0], 1)
        return I_x

    def round_optimum(self, x):
        """core/task/space.py:328-349."""
        x = np.array(x)
        if not ((x.ndim == 1) or (x.ndim == 2 and x.shape[0] == 1)):
            raise ValueError("Unexpected dimentionality of x. Got {}, expected (1, N) or (N,)".format(x.ndim))
        if x.ndim == 2:
            x = x[0]
        x_rounded = []
        for i, variable in enumerate(self.space_expanded):
            x_rounded.append(variable.round(x[i:i + 1]))
        return np.atleast_2d(np.concatenate(x_rounded))


def bounds_to_space(bounds):
    return [{'name': 'var_' + str(k + 1), 'type': 'continuous', 'domain': bounds[k], 'dimensionality': 1}
            for k in range(len(bounds))]


# ----------------------------------------------------------------------------------------------------------------------
# experiment design (random): np.random consumption order is part of the trajectory
# ----------------------------------------------------------------------------------------------------------------------
def samples_multidimensional_uniform(bounds, points_count):
    """random_design.py:80-91: ONE np.random.uniform(size=n) call per continuous dimension, in order."""
    dim = len(bounds)
    Z_rand = np.zeros(shape=(points_count, dim))
    for k in range(0, dim):
        Z_rand[:, k] = np.random.uniform(low=bounds[k][0], high=bounds[k][1], size=points_count)
    return Z_rand


class RandomDesign(object):
    def __init__(self, space):
        self.space = space

    def get_samples(self, init_points_count):
        if self.space.has_constraints():
            return self.get_samples_with_constraints(init_points_count)
        return self.get_samples_without_constraints(init_points_count)

    def get_samples_with_constraints(self, init_points_count):
        samples = np.empty((0, self.space.dimensionality))
        while samples.shape[0] < init_points_count:
            domain_samples = self.get_samples_without_constraints(init_points_count)
            valid_indices = (self.space.indicator_constraints(domain_samples) == 1).flatten()
            if sum(valid_indices) > 0:
                samples = np.vstack((samples, domain_samples[valid_indices, :]))
        return samples[0:init_points_count, :]

    def get_samples_without_constraints(self, init_points_count):
        samples = np.empty((init_points_count, self.space.dimensionality))
        for idx, var in enumerate(self.space.space_expanded):          # discrete variables first (random_design.py:43-46)
            if var.is_discrete():
                samples[:, idx] = np.atleast_2d(np.random.choice(var.domain, init_points_count)).flatten()
        if self.space.has_continuous():
            X_design = samples_multidimensional_uniform(self.space.get_continuous_bounds(), init_points_count)
            samples[:, self.space.get_continuous_dims()] = X_design
        return samples


def initial_design(design_name, space, init_points_count):
    if design_name != 'random':
        raise ValueError("only the 'random' design is provided (latin / sobol need pyDOE / sobol_seq): " + str(design_name))
    return RandomDesign(space).get_samples(init_points_count)


# ----------------------------------------------------------------------------------------------------------------------
# objective
# ----------------------------------------------------------------------------------------------------------------------
class SingleObjective(object):
    """core/task/objective.py:24-76 (sequential evaluation)."""

    def __init__(self, func, num_cores=1, objective_name='no_name', batch_type='synchronous', space=None):
        self.func, self.n_procs, self.num_evaluations, self.space, self.objective_name = func, num_cores, 0, space, objective_name

    def evaluate(self, x):
        cost_evals = []
        f_evals = np.empty(shape=[0, 1])
        for i in range(x.shape[0]):
            st_time = time.time()
            rlt = self.func(np.atleast_2d(x[i]))
            f_evals = np.vstack([f_evals, rlt])
            cost_evals += [time.time() - st_time]
        return f_evals, cost_evals


# ----------------------------------------------------------------------------------------------------------------------
# models
# ----------------------------------------------------------------------------------------------------------------------
class BOModel(object):
    """models/base.py:7-33."""
    MCMC_sampler = False
    analytical_gradient_prediction = False
    batched_rows_bitwise = False     # True: row i of a batched predict / acquisition call equals the single-row call bit for bit

    def updateModel(self, X_all, Y_all, X_new, Y_new):
        raise NotImplementedError

    def predict(self, X):
        raise NotImplementedError

    def predict_withGradients(self, X):
        return

    def get_fmin(self):
        raise NotImplementedError


class GPModel(BOModel):
    """models/gpmodel.py:9-177 on the B200 GPRegression."""

    analytical_gradient_prediction = True
    batched_rows_bitwise = True      # the device path's per-candidate sums do not depend on how many candidates share a call
    CONCURRENT_MIN_N, CONCURRENT_MAX_N = 384, 6144     # where several restarts at a time pay (scripts/concurrent_restarts_perf.py)

    def __init__(self, kernel=None, noise_var=None, exact_feval=False, optimizer='bfgs', max_iters=1000, optimize_restarts=5,
                 sparse=False, num_inducing=10, verbose=True, ARD=False, Gower=False, space=None, distributed_restarts=False,
                 concurrent_restarts=0):
        if sparse:
            raise NotImplementedError("sparse GPs are outside the B200 hot path (exact N x N on one GPU)")
        self.Gower, self.space = Gower, space
        self.kernel, self.noise_var, self.exact_feval = kernel, noise_var, exact_feval
        self.optimize_restarts, self.optimizer, self.max_iters, self.verbose = optimize_restarts, optimizer, max_iters, verbose
        self.sparse, self.num_inducing, self.model, self.ARD = sparse, num_inducing, None, ARD
        self._fmin = None
        # one optimize_restarts restart per torch.distributed rank (every rank must hold the same data and RNG state)
        self.distributed_restarts = distributed_restarts
        # > 1: that many optimize_restarts restarts at a time on this GPU (host threads, one model copy and CUDA stream each);
        # pays while one evaluation cannot fill the GPU and costs more than the thread set-up: CONCURRENT_MIN_N .. CONCURRENT_MAX_N points
        self.concurrent_restarts = concurrent_restarts

    @staticmethod
    def fromConfig(config):
        return GPModel(**config)

    def _create_model(self, X, Y):
        """gpmodel.py:50-76."""
        self.input_dim = X.shape[1]
        if self.kernel is None:
            kern = _kern.Matern52(self.input_dim, variance=1., ARD=self.ARD, Gower=self.Gower, space=self.space)
        else:
            kern = self.kernel
            self.kernel = None
        noise_var = Y.var() * 0.01 if self.noise_var is None else self.noise_var
        self.model = GPRegression(X, Y, kernel=kern, noise_var=noise_var)
        if self.exact_feval:
            self.model.Gaussian_noise.constrain_fixed(1e-6, warning=False)
        else:
            self.model.Gaussian_noise.constrain_bounded(1e-9, 1e6, warning=False)

    def updateModel(self, X_all, Y_all, X_new, Y_new):
        """gpmodel.py:78-93."""
        self._fmin = None
        if self.model is None:
            self._create_model(X_all, Y_all)
        else:
            self.model.set_XY(X_all, Y_all)
        if self.max_iters > 0:
            if self.optimize_restarts == 1:
                self.model.optimize(optimizer=self.optimizer, max_iters=self.max_iters, messages=False, ipython_notebook=False)
            else:
                extra = {"distributed": True} if self.distributed_restarts else {}
                if (self.concurrent_restarts > 1 and self.CONCURRENT_MIN_N <= self.model.X.shape[0] <= self.CONCURRENT_MAX_N
                        and not self.distributed_restarts):
                    extra = {"concurrent": self.concurrent_restarts}
                self.model.optimize_restarts(num_restarts=self.optimize_restarts, optimizer=self.optimizer,
                                             max_iters=self.max_iters, verbose=self.verbose, **extra)
        self._fmin = None

    def _nat(self):
        return self.model.posterior._nat

    def _predict(self, X, full_cov, include_likelihood):
        if X.ndim == 1:
            X = X[None, :]
        m, v = self.model.predict(X, full_cov=full_cov, include_likelihood=include_likelihood)
        v = np.clip(v, 1e-10, np.inf)
        return m, v

    def predict(self, X, with_noise=True):
        """gpmodel.py:102-112 -> (mean, standard deviation)."""
        m, v = self._predict(X, False, with_noise)
        return m, np.sqrt(v)

    def predict_covariance(self, X, with_noise=True):
        _, v = self._predict(X, True, with_noise)
        return v

    def get_fmin(self):
        """gpmodel.py:125-129.  The reference recomputes predict(model.X)[0].min() on every acquisition call; the value only
        changes with the model, so it is computed once per model state (on the device)."""
        if self._fmin is None or self._fmin[0] is not self.model.posterior:
            self._fmin = (self.model.posterior, self.model.posterior.fmin())
        return self._fmin[1]

    def predict_withGradients(self, X):
        """gpmodel.py:131-142."""
        if X.ndim == 1:
            X = X[None, :]
        r = self.model.posterior.acquisition("LCB", 0.0, 0.0, X, with_gradients=True, want_moments=True)
        return r["m"], r["s"], r["dmdx"], r["dsdx"]

    def acquisition_native(self, acq, par, X, with_gradients):
        """Fused device path of AcquisitionEI/LCB._compute_acq(_withGradients): returns (-f) [, (-df)] un-negated, i.e. the
        acquisition value and its gradient as the reference's _compute_acq* do."""
        if X.ndim == 1:
            X = X[None, :]
        r = self.model.posterior.acquisition(acq, par, self.get_fmin() if acq == "EI" else 0.0, X, with_gradients=with_gradients)
        if with_gradients:
            return -r["f"], -r["df"]
        return -r["f"]

    def copy(self):
        """gpmodel.py:144-159."""
        copied_model = GPModel(kernel=self.model.kern.copy(), noise_var=self.noise_var, exact_feval=self.exact_feval,
                               optimizer=self.optimizer, max_iters=self.max_iters, optimize_restarts=self.optimize_restarts,
                               verbose=self.verbose, ARD=self.ARD)
        copied_model._create_model(self.model.X, self.model.Y)
        copied_model.updateModel(self.model.X, self.model.Y, None, None)
        return copied_model

    def get_model_parameters(self):
        return np.atleast_2d(self.model[:])

    def get_model_parameters_names(self):
        return self.model.parameter_names_flat(include_fixed=True).tolist()

    def get_covariance_between_points(self, x1, x2):
        return self.model.posterior_covariance_between_points(x1, x2)


# ----------------------------------------------------------------------------------------------------------------------
# acquisitions
# ----------------------------------------------------------------------------------------------------------------------
class AcquisitionBase(object):
    """acquisitions/base.py:6-68."""
    analytical_gradient_prediction = False

    def __init__(self, model, space, optimizer, cost_withGradients=None):
        self.model, self.space, self.optimizer = model, space, optimizer
        self.analytical_gradient_acq = self.analytical_gradient_prediction and self.model.analytical_gradient_prediction
        self.cost_withGradients = constant_cost_withGradients if cost_withGradients is None else cost_withGradients

    def acquisition_function(self, x):
        f_acqu = self._compute_acq(x)
        cost_x, _ = self.cost_withGradients(x)
        return -(f_acqu * self.space.indicator_constraints(x)) / cost_x

    def acquisition_function_withGradients(self, x):
        f_acqu, df_acqu = self._compute_acq_withGradients(x)
        cost_x, cost_grad_x = self.cost_withGradients(x)
        f_acq_cost = f_acqu / cost_x
        df_acq_cost = (df_acqu * cost_x - f_acqu * cost_grad_x) / (cost_x ** 2)
        ind = self.space.indicator_constraints(x)
        return -f_acq_cost * ind, -df_acq_cost * ind

    @property
    def batched_rows_bitwise(self):
        """What lets AcquisitionOptimizer refine its anchors in lockstep (LockstepEvaluator): inherited from the model."""
        return bool(getattr(self.model, 'batched_rows_bitwise', False))

    def optimize(self, duplicate_manager=None):
        if not self.analytical_gradient_acq:
            out = self.optimizer.optimize(f=self.acquisition_function, duplicate_manager=duplicate_manager)
        else:
            out = self.optimizer.optimize(f=self.acquisition_function, f_df=self.acquisition_function_withGradients,
                                          duplicate_manager=duplicate_manager)
        return out

    def _compute_acq(self, x):
        raise NotImplementedError('')

    def _compute_acq_withGradients(self, x):
        raise NotImplementedError('')


class AcquisitionEI(AcquisitionBase):
    """acquisitions/EI.py:7-51."""
    analytical_gradient_prediction = True

    def __init__(self, model, space, optimizer=None, cost_withGradients=None, jitter=0.01):
        self.optimizer = optimizer
        super(AcquisitionEI, self).__init__(model, space, optimizer, cost_withGradients=cost_withGradients)
        self.jitter = jitter

    @staticmethod
    def fromConfig(model, space, optimizer, cost_withGradients, config):
        return AcquisitionEI(model, space, optimizer, cost_withGradients, jitter=config['jitter'])

    def _compute_acq(self, x):
        if hasattr(self.model, "acquisition_native"):
            return self.model.acquisition_native("EI", self.jitter, np.atleast_2d(x), False)
        m, s = self.model.predict(x)
        fmin = self.model.get_fmin()
        phi, Phi, u = get_quantiles(self.jitter, fmin, m, s)
        return s * (u * Phi + phi)

    def _compute_acq_withGradients(self, x):
        if hasattr(self.model, "acquisition_native"):
            return self.model.acquisition_native("EI", self.jitter, np.atleast_2d(x), True)
        fmin = self.model.get_fmin()
        m, s, dmdx, dsdx = self.model.predict_withGradients(x)
        phi, Phi, u = get_quantiles(self.jitter, fmin, m, s)
        return s * (u * Phi + phi), dsdx * phi - Phi * dmdx


class AcquisitionLCB(AcquisitionBase):
    """acquisitions/LCB.py:7-46."""
    analytical_gradient_prediction = True

    def __init__(self, model, space, optimizer=None, cost_withGradients=None, exploration_weight=2):
        self.optimizer = optimizer
        super(AcquisitionLCB, self).__init__(model, space, optimizer)
        self.exploration_weight = exploration_weight
        if cost_withGradients is not None:
            print('The set cost function is ignored! LCB acquisition does not make sense with cost.')

    def _compute_acq(self, x):
        if hasattr(self.model, "acquisition_native"):
            return self.model.acquisition_native("LCB", self.exploration_weight, np.atleast_2d(x), False)
        m, s = self.model.predict(x)
        return -m + self.exploration_weight * s

    def _compute_acq_withGradients(self, x):
        if hasattr(self.model, "acquisition_native"):
            return self.model.acquisition_native("LCB", self.exploration_weight, np.atleast_2d(x), True)
        m, s, dmdx, dsdx = self.model.predict_withGradients(x)
        return -m + self.exploration_weight * s, -dmdx + self.exploration_weight * dsdx


class AcquisitionLP(AcquisitionBase):
    """acquisitions/LP.py:10-140: local-penalisation wrapper used by the batch evaluator.  The penalised acquisition lives in
    log space: f(x) = -T(acq(x)) - sum_b log Phi((|x - x_b| - r_b) / s_b), T = log(. + 1e-50) or log softplus(.).

    With the B200 GPModel and an EI / LCB base acquisition (constant cost, no constraints) value and gradient come from one
    device pass (gpb_model_acquisition_lp); any other combination runs the reference's NumPy formulas on whatever the wrapped
    acquisition returns."""
    analytical_gradient_prediction = True

    def __init__(self, model, space, optimizer, acquisition, transform='none'):
        super(AcquisitionLP, self).__init__(model, space, optimizer)
        self.acq = acquisition
        self.transform = transform.lower()
        if isinstance(acquisition, AcquisitionLCB) and self.transform == 'none':
            self.transform = 'softplus'
        self.X_batch = None
        self.r_x0 = None
        self.s_x0 = None
        self._pushed_to = None

    @property
    def batched_rows_bitwise(self):
        """Only the device pass qualifies: the reference's NumPy gradient formula broadcasts correctly for one point at a time only
        (LP.py:129-132)."""
        return bool(getattr(self.model, 'batched_rows_bitwise', False)) and self._native_kind() is not None

    # -- device path -----------------------------------------------------------------------------------------------------
    def _native_kind(self):
        if not (isinstance(self.model, GPModel) and self.model.model is not None
                and hasattr(self.model.model.posterior, "acquisition_lp")):
            return None
        if self.transform not in ('none', 'softplus') or self.space.has_constraints():
            return None
        if getattr(self.acq, "cost_withGradients", None) is not constant_cost_withGradients:
            return None
        if type(self.acq) is AcquisitionEI:
            return "EI", self.acq.jitter
        if type(self.acq) is AcquisitionLCB:
            return "LCB", self.acq.exploration_weight
        return None

    def _push(self, post):
        """Upload the batch points / hammer parameters when they or the resident model changed."""
        key = (id(getattr(post, "_nat", post)), id(self.X_batch), self.transform)
        if self._pushed_to != key:
            post.set_penalizers(self.transform, self.X_batch, self.r_x0, self.s_x0)
            self._pushed_to = key

    def _native_call(self, x, with_gradients):
        kind, par = self._native_kind()
        post = self.model.model.posterior
        self._push(post)
        fmin = self.model.get_fmin() if kind == "EI" else 0.0
        return post.acquisition_lp(kind, par, fmin, np.atleast_2d(x), with_gradients=with_gradients)

    # -- reference logic -------------------------------------------------------------------------------------------------
    def update_batches(self, X_batch, L, Min):
        """LP.py:40-47."""
        self.X_batch = X_batch
        self._pushed_to = None
        if X_batch is not None:
            self.r_x0, self.s_x0 = self._hammer_function_precompute(X_batch, L, Min, self.model)

    def _hammer_function_precompute(self, x0, L, Min, model):
        """LP.py:49-62 (note: the reference takes the square root of the predicted standard deviation once more)."""
        if x0 is None:
            return None, None
        if len(x0.shape) == 1:
            x0 = x0[None, :]
        m, sd = model.predict(x0)
        pred = sd.copy()
        pred[pred < 1e-16] = 1e-16
        s = np.sqrt(pred)
        r_x0 = (m - Min) / L
        s_x0 = s / L
        return r_x0.flatten(), s_x0.flatten()

    def _hammer_function(self, x, x0, r_x0, s_x0):
        """LP.py:64-68."""
        from scipy.stats import norm
        return norm.logcdf((np.sqrt((np.square(np.atleast_2d(x)[:, None, :] - np.atleast_2d(x0)[None, :, :])).sum(-1)) - r_x0) / s_x0)

    def _penalized_acquisition(self, x, model, X_batch, r_x0, s_x0):
        """LP.py:70-89."""
        fval = -self.acq.acquisition_function(x)[:, 0]
        if self.transform == 'softplus':
            fval_org = fval.copy()
            fval[fval_org >= 40.] = np.log(fval_org[fval_org >= 40.])
            fval[fval_org < 40.] = np.log(np.log1p(np.exp(fval_org[fval_org < 40.])))
        elif self.transform == 'none':
            fval = np.log(fval + 1e-50)
        fval = -fval
        if X_batch is not None:
            h_vals = self._hammer_function(x, X_batch, r_x0, s_x0)
            fval += -h_vals.sum(axis=-1)
        return fval

    def _d_hammer_function(self, x, X_batch, r_x0, s_x0):
        """LP.py:91-104 (the sum over the batch is returned as one scalar per point, broadcast over the dimensions)."""
        from scipy.stats import norm
        dx = np.atleast_2d(x)[:, None, :] - np.atleast_2d(X_batch)[None, :, :]
        nm = np.sqrt((np.square(dx)).sum(-1))
        z = (nm - r_x0) / s_x0
        h_func = norm.cdf(z)
        d = 1. / (s_x0 * np.sqrt(2 * np.pi) * h_func) * np.exp(-np.square(z) / 2) / nm
        d[h_func < 1e-50] = 0.
        d = d[:, :, None]
        return d.sum(axis=1)

    def acquisition_function(self, x):
        """LP.py:106-111."""
        if self._native_kind() is not None:
            return self._native_call(x, False)
        return self._penalized_acquisition(x, self.model, self.X_batch, self.r_x0, self.s_x0)

    def d_acquisition_function(self, x):
        """LP.py:113-132."""
        if self._native_kind() is not None:
            return self._native_call(x, True)[1]
        x = np.atleast_2d(x)
        if self.transform == 'softplus':
            fval = -self.acq.acquisition_function(x)[:, 0]
            scale = 1. / (np.log1p(np.exp(fval)) * (1. + np.exp(-fval)))
        elif self.transform == 'none':
            fval = -self.acq.acquisition_function(x)[:, 0]
            scale = 1. / fval
        else:
            scale = 1.
        _, grad_acq_x = self.acq.acquisition_function_withGradients(x)
        if self.X_batch is None:
            return scale * grad_acq_x
        return scale * grad_acq_x - self._d_hammer_function(x, self.X_batch, self.r_x0, self.s_x0)

    def acquisition_function_withGradients(self, x):
        """LP.py:134-140."""
        if self._native_kind() is not None:
            return self._native_call(x, True)
        return self.acquisition_function(x), self.d_acquisition_function(x)


# ----------------------------------------------------------------------------------------------------------------------
# acquisition optimiser
# ----------------------------------------------------------------------------------------------------------------------
class OptLbfgs(object):
    """optimization/optimizer.py:28-61."""

    def __init__(self, bounds, maxiter=1000):
        self.bounds, self.maxiter = bounds, maxiter

    def optimize(self, x0, f=None, df=None, f_df=None):
        import scipy.optimize
        if f_df is None and df is not None:
            f_df = lambda x: (float(f(x)), df(x))  # noqa: E731
        if f_df is not None:
            def _f_df(x):
                # the reference evaluates f(x) AND f_df(x) per step (optimizer.py:46-47); both come from the same device
                # kernels, so one fused call returns bit-identical values
                fx, dfx = f_df(x)
                return fx, dfx[0]
        if f_df is None and df is None:
            res = scipy.optimize.fmin_l_bfgs_b(f, x0=x0, bounds=self.bounds, approx_grad=True, maxiter=self.maxiter)
        else:
            res = scipy.optimize.fmin_l_bfgs_b(_f_df, x0=x0, bounds=self.bounds, maxiter=self.maxiter)
        if res[2]['task'] == b'ABNORMAL_TERMINATION_IN_LNSRCH' or res[2]['task'] == 'ABNORMAL_TERMINATION_IN_LNSRCH':
            result_x = np.atleast_2d(x0)
            result_fx = np.atleast_2d(f(x0))
        else:
            result_x = np.atleast_2d(res[0])
            result_fx = np.atleast_2d(res[1])
        return result_x, result_fx


def apply_optimizer(optimizer, x0, f=None, df=None, f_df=None, duplicate_manager=None, context_manager=None, space=None):
    """optimization/optimizer.py:130-168 without context variables."""
    x0 = np.atleast_2d(x0)
    if duplicate_manager and duplicate_manager.is_unzipped_x_duplicate(x0):
        raise ValueError("The starting point of the optimizer cannot be a duplicate.")

    # OptimizationWithContext.f_nc / f_df_nc (optimizer.py:200-232): SciPy hands the objective a 1-D x; the reference always
    # routes it through these wrappers (a ContextManager exists even without context), so acquisitions only ever see 2-D input
    def f_nc(x):
        x = np.atleast_2d(x)
        return f(x)[0] if x.shape[0] == 1 else f(x)

    f_df_nc = None if f_df is None else (lambda x: f_df(np.atleast_2d(x)))
    optimized_x, _ = optimizer.optimize(x0, f_nc, df, f_df_nc)
    suggested_x_rounded = space.round_optimum(optimized_x)
    if duplicate_manager and duplicate_manager.is_unzipped_x_duplicate(suggested_x_rounded):
        return x0, np.atleast_2d(f(x0))            # optimizer.py:163-164: fall back to the (non-duplicate) anchor
    return suggested_x_rounded, f(suggested_x_rounded)


class LockstepEvaluator(object):
    """Coalesces the f_df requests of several concurrent L-BFGS-B runs into ONE device call.

    The refinements of the anchor points (acquisition_optimizer.py:60-75) are independent; each one is a chain of several hundred
    M = 1 acquisition calls, and one such call streams the whole triangle of L^-1 twice (2 x 8 N^2 / 2 bytes).  The device path
    serves up to 8 candidates for the SAME single pass over the triangle, with per-candidate sums that are bit-identical to the
    M = 1 call (fixed reduction orders that do not depend on the number of candidates).  So the runs are advanced in lockstep: every
    run lives in its own host thread with an UNMODIFIED scipy.optimize.fmin_l_bfgs_b; a run that needs (f, df) parks its x here, and
    when every still-active run has parked one, a single batched call answers all of them.  Each run sees exactly the values the
    sequential loop would have given it -> same trajectories, same optimum, 1 / (number of runs) of the bytes streamed.
    """

    def __init__(self, f_df, n_workers, max_rows=8):
        import threading
        self.f_df, self.max_rows = f_df, max_rows
        self.cv = threading.Condition()
        self.active = n_workers
        self.pending = {}      # worker -> x (1, d)
        self.results = {}      # worker -> (f (1, 1), df (1, d)) or an exception
        self.calls = 0         # device calls issued
        self.requests = 0      # f_df requests answered

    def _flush_locked(self):
        """Called with the lock held when every active run has a request parked: answer them, max_rows per device call."""
        workers = sorted(self.pending)
        for a in range(0, len(workers), self.max_rows):
            grp = workers[a:a + self.max_rows]
            X = np.vstack([self.pending[w] for w in grp])
            try:
                fx, dfx = self.f_df(X)
                fx, dfx = np.asarray(fx), np.asarray(dfx)          # row slices keep the shapes a single-row call returns
                assert fx.shape[0] == len(grp) and dfx.shape[0] == len(grp)
                for i, w in enumerate(grp):
                    self.results[w] = (fx[i:i + 1].copy(), dfx[i:i + 1].copy())
            except Exception as exc:                       # every run of the group sees the failure
                for w in grp:
                    self.results[w] = exc
            self.calls += 1
            self.requests += len(grp)
        self.pending.clear()
        self.cv.notify_all()

    def evaluate(self, worker, x):
        with self.cv:
            self.pending[worker] = np.atleast_2d(np.array(x, dtype=np.float64))
            if len(self.pending) == self.active:
                self._flush_locked()
            while worker not in self.results:
                self.cv.wait()
            r = self.results.pop(worker)
        if isinstance(r, Exception):
            raise r
        return r

    def retire(self, worker):
        """The run of `worker` has ended: the others no longer wait for it."""
        with self.cv:
            self.active -= 1
            if self.pending and len(self.pending) == self.active:
                self._flush_locked()


class ObjectiveAnchorPointsGenerator(object):
    """optimization/anchor_points_generator.py:8-98: 1000 random points, scored in ONE batched call, keep the 5 lowest."""

    def __init__(self, space, design_type, objective, num_samples=1000, world=1, rank=0):
        self.space, self.design_type, self.objective, self.num_samples = space, design_type, objective, num_samples
        self.world, self.rank = world, rank

    def get_anchor_point_scores(self, X):
        if self.world > 1 and X.shape[0] >= self.world:
            return self._sharded_scores(X)
        return self.objective(X).flatten()

    def _sharded_scores(self, X):
        """Every rank holds the same model and draws the same candidates; rank r scores rows [lo, hi) only and ONE all-gather of the
        scores gives every rank the vector the single-process call computes (a candidate's score does not depend on which other rows
        share its call), so the argsort below -- and with it the whole BO trajectory -- is unchanged.  At N = 32768 the reference's
        1000-candidate scoring call costs 60 ms per batch element; sharded over 8 GPUs 8 ms."""
        import torch
        import torch.distributed as dist
        from .sharded import divide_candidates
        n = X.shape[0]
        lo, hi = divide_candidates(n, self.rank, self.world)
        mine = np.asarray(self.objective(X[lo:hi]), dtype=np.float64).flatten()
        width = -(-n // self.world)
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        buf = torch.full((width,), float("nan"), dtype=torch.float64)
        buf[:hi - lo] = torch.from_numpy(mine)
        buf = buf.to(dev)
        out = [torch.empty_like(buf) for _ in range(self.world)]
        dist.all_gather(out, buf)
        parts = []
        for r in range(self.world):
            a, b = divide_candidates(n, r, self.world)
            parts.append(out[r][:b - a].cpu().numpy())
        return np.concatenate(parts)

    def get(self, num_anchor=5, duplicate_manager=None, unique=False, context_manager=None):
        X = initial_design(self.design_type, self.space, self.num_samples)
        if unique:
            X = np.vstack(sorted(list({tuple(x) for x in X})))
        X = self.space.unzip_inputs(X)
        if duplicate_manager:                                          # anchor_points_generator.py:44-58
            keep = [i for i, x in enumerate(X) if not duplicate_manager.is_unzipped_x_duplicate(x)]
            if not keep:
                raise FullyExploredOptimizationDomainError("No anchor points could be generated ({} used samples, {} requested "
                                                           "anchor points).".format(self.num_samples, num_anchor))
            if len(keep) < num_anchor:
                print("Warning: expecting {} anchor points, only {} available.".format(num_anchor, len(keep)))
            X = X[keep, :]
        scores = self.get_anchor_point_scores(X)
        # np.argsort's default introsort is not stable; ties are resolved towards the lowest index here (and in the oracle)
        return X[np.argsort(scores, kind='stable')[:min(len(scores), num_anchor)], :]


class AcquisitionOptimizer(object):
    """optimization/acquisition_optimizer.py:16-77."""

    def __init__(self, space, optimizer='lbfgs', **kwargs):
        if optimizer != 'lbfgs':
            raise NotImplementedError("only 'lbfgs' is provided (DIRECT / CMA need external packages)")
        self.space, self.optimizer_name, self.kwargs = space, optimizer, kwargs
        if 'model' in self.kwargs:
            self.model = self.kwargs['model']
        self.context_manager = None

    def optimize(self, f=None, df=None, f_df=None, duplicate_manager=None):
        self.f, self.df, self.f_df = f, df, f_df
        self.optimizer = OptLbfgs(self.space.get_bounds())
        world, rank = self._world()
        anchor_points = ObjectiveAnchorPointsGenerator(self.space, 'random', f, world=world, rank=rank).get(duplicate_manager=duplicate_manager)
        run = lambda a: apply_optimizer(self.optimizer, a, f=f, df=None, f_df=f_df, duplicate_manager=duplicate_manager,  # noqa: E731
                                        space=self.space)
        if world > 1:
            optimized_points = self._optimize_anchors_distributed(anchor_points, run, world, rank, f, f_df, duplicate_manager)
        elif self._lockstep_ok(f_df, anchor_points):
            optimized_points = self._optimize_anchors_lockstep(anchor_points, f, f_df, duplicate_manager)
        else:
            optimized_points = [run(a) for a in anchor_points]
        x_min, fx_min = min(optimized_points, key=lambda t: t[1])
        return x_min, fx_min

    # -- all anchors refined concurrently, their M = 1 requests coalesced into one M <= 8 device call (LockstepEvaluator) ----------
    # Needs an f_df whose rows are bit-identical to single-row calls: the CUDA acquisition path declares it
    # (`batched_rows_bitwise`); the kwarg lockstep_anchors=False forces the sequential loop.
    LOCKSTEP_MIN_N = 1024      # below this a device call is tens of microseconds and the threads' hand-overs cost more than they save

    def _lockstep_ok(self, f_df, anchor_points):
        """kwarg lockstep_anchors: True forces it, False forbids it, absent = automatic (models of at least LOCKSTEP_MIN_N points:
        BASELINE config 1, N <= 35, ran 10% slower in lock step -- 0.57 s against 0.51 s for the 30 iterations)."""
        mode = self.kwargs.get('lockstep_anchors', 'auto')
        if f_df is None or len(anchor_points) < 2 or mode is False:
            return False
        owner = getattr(f_df, '__self__', None)
        if not getattr(owner, 'batched_rows_bitwise', False):
            return False
        if mode is True:
            return True
        gp = getattr(getattr(owner, 'model', None), 'model', None)
        n = getattr(getattr(gp, 'X', None), 'shape', (0,))[0]
        return n >= self.LOCKSTEP_MIN_N

    def _optimize_anchors_lockstep(self, anchor_points, f, f_df, duplicate_manager):
        import threading
        anchor_points = [np.atleast_2d(a) for a in anchor_points]
        # the duplicate check of apply_optimizer happens before any evaluation: keep its error in anchor order
        if duplicate_manager:
            for a in anchor_points:
                if duplicate_manager.is_unzipped_x_duplicate(a):
                    raise ValueError("The starting point of the optimizer cannot be a duplicate.")
        na = len(anchor_points)
        ev = LockstepEvaluator(f_df, na)
        out, errors = [None] * na, [None] * na
        dev = None
        try:
            import torch
            if torch.cuda.is_available() and torch.cuda.is_initialized():
                dev = torch.cuda.current_device()
        except ImportError:
            torch = None

        def work(i):
            try:
                if dev is not None:
                    torch.cuda.set_device(dev)              # the current CUDA device is per host thread
                out[i] = apply_optimizer(self.optimizer, anchor_points[i], f=f, df=None, f_df=lambda x: ev.evaluate(i, x),
                                         duplicate_manager=duplicate_manager, space=self.space)
            except Exception as exc:
                errors[i] = exc
            finally:
                ev.retire(i)

        threads = [threading.Thread(target=work, args=(i,)) for i in range(na)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        for e in errors:
            if e is not None:
                raise e
        self.lockstep_stats = {"device_calls": ev.calls, "requests": ev.requests}
        return out

    # -- one L-BFGS-B refinement per torch.distributed rank (kwarg distributed_anchors=True) ---------------------------------
    # The refinements of the anchor points (acquisition_optimizer.py:68-72) are independent and deterministic, so rank r runs
    # anchors r, r + world, ...; one all-reduce of an (anchors x (1 + d)) table (every row written by exactly one rank) gives
    # every rank the list the sequential loop would have produced, in the same order: same minimum, same tie-breaking.  Every
    # rank must hold the same model and the same NumPy RNG state (the anchors are drawn on every rank).
    def _world(self):
        if not self.kwargs.get('distributed_anchors', False):
            return 1, 0
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            return 1, 0
        return dist.get_world_size(), dist.get_rank()

    def _optimize_anchors_distributed(self, anchor_points, run, world, rank, f=None, f_df=None, duplicate_manager=None):
        import torch
        import torch.distributed as dist
        anchor_points = np.atleast_2d(anchor_points)
        na, d = anchor_points.shape
        table = np.full((na, 1 + d), np.inf)
        mine = list(range(rank, na, world))
        if len(mine) > 1 and self._lockstep_ok(f_df, anchor_points[mine]):
            # fewer ranks than anchors: this rank's anchors advance in lock step (one coalesced device call per joint step)
            results = self._optimize_anchors_lockstep(anchor_points[mine], f, f_df, duplicate_manager)
        else:
            results = [run(anchor_points[i]) for i in mine]
        for i, (x, fx) in zip(mine, results):
            table[i, 0], table[i, 1:] = float(np.asarray(fx).ravel()[0]), np.asarray(x).ravel()
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        t = torch.from_numpy(table).to(dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        table = t.cpu().numpy()
        return [(table[i, 1:][None, :].copy(), np.array([[table[i, 0]]])) for i in range(na)]


class Sequential(object):
    """core/evaluators/sequential.py:7-23."""

    def __init__(self, acquisition, batch_size=1):
        self.acquisition, self.batch_size = acquisition, batch_size

    def compute_batch(self, duplicate_manager=None, context_manager=None):
        x, _ = self.acquisition.optimize(duplicate_manager=duplicate_manager)
        return x


def estimate_L(model, bounds, storehistory=True):
    """core/evaluators/batch_local_penalization.py:50-70: Lipschitz constant = max over the domain of |d mu / dx|.
    500 uniform samples (one np.random.uniform call per dimension, util/general.py:63-73) + the training inputs are scored
    in one batched call, then bounded L-BFGS-B (maxiter 200, finite-difference gradient like the reference) from the best."""
    import scipy.optimize

    def df(x, model, x0):
        x = np.atleast_2d(x)
        try:
            dmdx, _ = model.predictive_gradients(x, want_var=False)   # the reference computes and discards the variance part
        except TypeError:
            dmdx, _ = model.predictive_gradients(x)
        res = np.sqrt((dmdx * dmdx).sum(1))
        return -res

    samples = samples_multidimensional_uniform(bounds, 500)
    samples = np.vstack([samples, model.X])
    pred_samples = df(samples, model, 0)
    x0 = samples[np.argmin(pred_samples)]
    res = scipy.optimize.minimize(lambda x, *a: float(df(x, *a)[0, 0]), x0, method='L-BFGS-B', bounds=bounds, args=(model, x0),
                                  options={'maxiter': 200})
    minusL = float(np.atleast_2d(res.fun)[0][0])
    L = -minusL
    if L < 1e-7:
        L = 10  # to avoid problems in cases in which the model is flat
    return L


class LocalPenalization(object):
    """core/evaluators/batch_local_penalization.py:9-48: batch = first point from the plain acquisition, the others from the
    acquisition penalised around the points already in the batch."""

    def __init__(self, acquisition, batch_size):
        self.acquisition, self.batch_size = acquisition, batch_size

    def compute_batch(self, duplicate_manager=None, context_manager=None):
        assert isinstance(self.acquisition, AcquisitionLP)
        self.acquisition.update_batches(None, None, None)
        X_batch = self.acquisition.optimize()[0]
        k = 1
        if self.batch_size > 1:
            L = estimate_L(self.acquisition.model.model, self.acquisition.space.get_bounds())
            Min = self.acquisition.model.model.Y.min()
        while k < self.batch_size:
            self.acquisition.update_batches(X_batch, L, Min)
            new_sample = self.acquisition.optimize()[0]
            X_batch = np.vstack((X_batch, new_sample))
            k += 1
        self.acquisition.update_batches(None, None, None)
        return X_batch


# ----------------------------------------------------------------------------------------------------------------------
# BO loop
# ----------------------------------------------------------------------------------------------------------------------
class BO(object):
    """core/bo.py:20-260."""

    def __init__(self, model, space, objective, acquisition, evaluator, X_init, Y_init=None, cost=None, normalize_Y=True,
                 model_update_interval=1, de_duplication=False):
        self.model, self.space, self.objective, self.acquisition, self.evaluator = model, space, objective, acquisition, evaluator
        self.normalize_Y, self.model_update_interval = normalize_Y, model_update_interval
        self.X, self.Y = X_init, Y_init
        self.normalization_type = 'stats'
        self.de_duplication = de_duplication
        self.model_parameters_iterations = None
        self.context = None
        self.num_acquisitions = 0

    def suggest_next_locations(self, context=None, pending_X=None, ignored_X=None):
        self.model_parameters_iterations = None
        self.num_acquisitions = 0
        self.context = context
        self._update_model(self.normalization_type)
        return self._compute_next_evaluations(pending_zipped_X=pending_X, ignored_zipped_X=ignored_X)

    def run_optimization(self, max_iter=0, max_time=np.inf, eps=1e-8, context=None, verbosity=False,
                         save_models_parameters=True, report_file=None, evaluations_file=None, models_file=None):
        if self.objective is None:
            raise InvalidConfigError("Cannot run the optimization loop without the objective function")
        self.verbosity = verbosity
        self.save_models_parameters = save_models_parameters
        self.model_parameters_iterations = None
        self.context = context
        self.eps = eps
        if (max_iter is None) and (max_time is None):
            self.max_iter, self.max_time = 0, np.inf
        elif (max_iter is None) and (max_time is not None):
            self.max_iter, self.max_time = np.inf, max_time
        elif (max_iter is not None) and (max_time is None):
            self.max_iter, self.max_time = max_iter, np.inf
        else:
            self.max_iter, self.max_time = max_iter, max_time
        if self.X is not None and self.Y is None:
            self.Y, _ = self.objective.evaluate(self.X)
        self.time_zero = time.time()
        self.cum_time = 0
        self.num_acquisitions = 0
        self.suggested_sample = self.X
        self.Y_new = self.Y
        while self.max_time > self.cum_time:
            try:
                self._update_model(self.normalization_type)
            except np.linalg.LinAlgError:
                break
            if (self.num_acquisitions >= self.max_iter
                    or (len(self.X) > 1 and self._distance_last_evaluations() <= self.eps)):
                break
            self.suggested_sample = self._compute_next_evaluations()
            self.X = np.vstack((self.X, self.suggested_sample))
            self.evaluate_objective()
            self.cum_time = time.time() - self.time_zero
            self.num_acquisitions += 1
            if verbosity:
                print("num acquisition: {}, time elapsed: {:.2f}s".format(self.num_acquisitions, self.cum_time))
        self._compute_results()

    def evaluate_objective(self):
        self.Y_new, _ = self.objective.evaluate(self.suggested_sample)
        self.Y = np.vstack((self.Y, self.Y_new))

    def _compute_results(self):
        self.Y_best = best_value(self.Y)
        self.x_opt = self.X[np.argmin(self.Y), :]
        self.fx_opt = np.min(self.Y)

    def _distance_last_evaluations(self):
        if self.X.shape[0] < 2:
            return np.inf
        return np.sqrt(np.sum((self.X[-1, :] - self.X[-2, :]) ** 2))

    def _compute_next_evaluations(self, pending_zipped_X=None, ignored_zipped_X=None):
        """core/bo.py:216-234."""
        duplicate_manager = None
        if self.de_duplication:
            duplicate_manager = DuplicateManager(space=self.space, zipped_X=self.X, pending_zipped_X=pending_zipped_X,
                                                 ignored_zipped_X=ignored_zipped_X)
        return self.space.zip_inputs(self.evaluator.compute_batch(duplicate_manager=duplicate_manager, context_manager=None))

    def _update_model(self, normalization_type='stats'):
        if self.num_acquisitions % self.model_update_interval == 0:
            X_inmodel = self.space.unzip_inputs(self.X)
            Y_inmodel = normalize(self.Y, normalization_type) if self.normalize_Y else self.Y
            self.model.updateModel(X_inmodel, Y_inmodel, None, None)
        self._save_model_parameter_values()

    def _save_model_parameter_values(self):
        if self.model_parameters_iterations is None:
            self.model_parameters_iterations = self.model.get_model_parameters()
        else:
            self.model_parameters_iterations = np.vstack((self.model_parameters_iterations, self.model.get_model_parameters()))

    def get_evaluations(self):
        return self.X.copy(), self.Y.copy()


class BayesianOptimization(BO):
    """methods/bayesian_optimization.py:76-202 (GP model, EI / LCB acquisition, L-BFGS-B acquisition optimiser)."""

    def __init__(self, f, domain=None, constraints=None, cost_withGradients=None, model_type='GP', X=None, Y=None,
                 initial_design_numdata=5, initial_design_type='random', acquisition_type='EI', normalize_Y=True,
                 exact_feval=False, acquisition_optimizer_type='lbfgs', model_update_interval=1, evaluator_type='sequential',
                 batch_size=1, num_cores=1, verbosity=False, verbosity_model=False, maximize=False, de_duplication=False, **kwargs):
        self.modular_optimization = False
        self.initial_iter = True
        self.verbosity, self.verbosity_model = verbosity, verbosity_model
        self.model_update_interval, self.de_duplication, self.kwargs = model_update_interval, de_duplication, kwargs
        self.constraints, self.domain = constraints, domain
        self.space = Design_space(self.domain, self.constraints)
        self.maximize = maximize
        self.objective_name = kwargs.get('objective_name', 'no_name')
        self.batch_size, self.num_cores = batch_size, num_cores
        if f is not None:
            self.f = self._sign(f)
            self.objective = SingleObjective(self.f, self.batch_size, self.objective_name)
        else:
            self.f, self.objective = None, None
        self.cost_withGradients = cost_withGradients
        self.X, self.Y = X, Y
        self.initial_design_type, self.initial_design_numdata = initial_design_type, initial_design_numdata
        self._init_design_chooser()
        self.model_type, self.exact_feval, self.normalize_Y = model_type, exact_feval, normalize_Y
        if 'model' in kwargs and isinstance(kwargs['model'], BOModel):
            self.model = kwargs['model']
            self.model_type = 'User defined model used.'
        else:
            self.model = self._model_chooser()
        self.acquisition_optimizer_type = acquisition_optimizer_type
        self.acquisition_optimizer = AcquisitionOptimizer(self.space, self.acquisition_optimizer_type, model=self.model,
                                                          distributed_anchors=kwargs.get('distributed_anchors', False))
        self.acquisition_type = acquisition_type
        if 'acquisition' in kwargs and isinstance(kwargs['acquisition'], AcquisitionBase):
            self.acquisition = kwargs['acquisition']
            self.acquisition_type = 'User defined acquisition used.'
        else:
            self.acquisition = self._acquisition_chooser()
        self.evaluator_type = evaluator_type
        self.evaluator = self._evaluator_chooser()
        super(BayesianOptimization, self).__init__(model=self.model, space=self.space, objective=self.objective,
                                                   acquisition=self.acquisition, evaluator=self.evaluator, X_init=self.X,
                                                   Y_init=self.Y, cost=None, normalize_Y=self.normalize_Y,
                                                   model_update_interval=self.model_update_interval,
                                                   de_duplication=self.de_duplication)

    def _model_chooser(self):
        """util/arguments_manager.py:78-110 (defaults: lbfgs, max_iters 1000, 5 restarts, ARD False)."""
        if self.model_type != 'GP':
            raise NotImplementedError("model_type %r is outside the B200 hot path" % (self.model_type,))
        kw = self.kwargs
        Gower = kw.get('Gower', False)
        return GPModel(kw.get('kernel', None), kw.get('noise_var', None), self.exact_feval, kw.get('model_optimizer_type', 'lbfgs'),
                       kw.get('max_iters', 1000), kw.get('optimize_restarts', 5), False, kw.get('num_inducing', 10),
                       kw.get('verbosity_model', False), kw.get('ARD', False), Gower, self.space if Gower is True else None)

    def _acquisition_chooser(self):
        """util/arguments_manager.py:42-75 (jitter 0.01, weight 2)."""
        jitter = self.kwargs.get('acquisition_jitter', 0.01)
        weight = self.kwargs.get('acquisition_weight', 2)
        if self.acquisition_type is None or self.acquisition_type == 'EI':
            return AcquisitionEI(self.model, self.space, self.acquisition_optimizer, self.cost_withGradients, jitter)
        if self.acquisition_type == 'LCB':
            return AcquisitionLCB(self.model, self.space, self.acquisition_optimizer, self.cost_withGradients, weight)
        raise Exception('Invalid acquisition selected.')

    def _evaluator_chooser(self):
        """util/arguments_manager.py:17-38."""
        if self.batch_size == 1 or self.evaluator_type == 'sequential':
            return Sequential(self.acquisition)
        if self.evaluator_type == 'local_penalization':
            if self.model_type not in ['GP', 'User defined model used.']:
                raise InvalidConfigError('local_penalization evaluator can only be used with GP models')
            acq = self.acquisition
            if not isinstance(acq, AcquisitionLP):
                acq = AcquisitionLP(self.model, self.space, self.acquisition_optimizer, self.acquisition,
                                    self.kwargs.get('acquisition_transformation', 'none'))
            return LocalPenalization(acq, self.batch_size)
        raise NotImplementedError("evaluator_type %r is outside the B200 hot path (random / Thompson batches)" % (self.evaluator_type,))

    def _init_design_chooser(self):
        if self.f is None and (self.X is None or self.Y is None):
            raise InvalidConfigError("Initial data for both X and Y is required when objective function is not provided")
        if self.X is None:
            self.X = initial_design(self.initial_design_type, self.space, self.initial_design_numdata)
            self.Y, _ = self.objective.evaluate(self.X)
        elif self.X is not None and self.Y is None:
            self.Y, _ = self.objective.evaluate(self.X)

    def _sign(self, f):
        if self.maximize:
            f_copy = f

            def f(x):
                return -f_copy(x)
        return f


class ModularBayesianOptimization(BO):
    """methods/modular_bayesian_optimization.py:6-40: the BO loop around handlers the caller built (model, space, objective,
    acquisition, evaluator, initial data) -- the second plug-in level of SURVEY 8b."""

    def __init__(self, model, space, objective, acquisition, evaluator, X_init, Y_init=None, cost=None, normalize_Y=True,
                 model_update_interval=1, de_duplication=False):
        self.initial_iter = True
        self.modular_optimization = True
        super(ModularBayesianOptimization, self).__init__(model=model, space=space, objective=objective, acquisition=acquisition,
                                                          evaluator=evaluator, X_init=X_init, Y_init=Y_init, cost=cost,
                                                          normalize_Y=normalize_Y, model_update_interval=model_update_interval,
                                                          de_duplication=de_duplication)
