"""Run the reference's OWN Bayesian-optimisation loop (GPyOpt 1.2.5 under /root/reference) in this container.

Container-only test infrastructure, like ref_harness.py.  On top of ref_harness.load(full_paramz=True) -- the reference's
numerical GPy/GPyOpt modules with the repo's paramz restatement as the parameter machinery -- this imports the reference's
BO-loop sources unmodified:
    GPyOpt/core/bo.py, core/task/{space,variables,objective,cost}.py, core/evaluators/{base,sequential,
    batch_local_penalization}.py, optimization/{acquisition_optimizer,optimizer,anchor_points_generator}.py,
    experiment_design/{base,random_design}.py, util/duplicate_manager.py
and assembles them the way methods/bayesian_optimization.py:76-171 does (that file itself drags in every model / acquisition /
design of GPyOpt through util/arguments_manager.py and cannot be imported here).  Only two pieces of glue are restated:
the `initial_design` dispatcher of experiment_design/__init__.py:8-24 (the 'random' branch) and the constructor wiring.

What this pins: the repo's host mirror of the BO loop (gaussian_process_optimization_b200/gpyopt.py) -- initial design and RNG
consumption order, Y normalisation, model update with restarts, anchor-point generation and selection, L-BFGS-B refinement,
rounding, stopping rule -- against the reference's code, evaluation by evaluation.
"""
import importlib
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as rh  # noqa: E402


def load_bo():
    ns = rh.load(full_paramz=True)
    o = os.path.join(rh.REF, "GPyOpt", "GPyOpt")
    for name, sub in (("GPyOpt.optimization", "optimization"), ("GPyOpt.experiment_design", "experiment_design"),
                      ("GPyOpt.core.evaluators", os.path.join("core", "evaluators")), ("GPyOpt.plotting", "plotting")):
        if name not in sys.modules:
            rh._pkg(name, os.path.join(o, sub))
    sys.modules["GPyOpt.models"].base = sys.modules["GPyOpt.models.base"]
    # experiment designs: the reference's RandomDesign; latin / sobol / grid need pyDOE / sobol_seq and are not on the path
    importlib.import_module("GPyOpt.core.task.variables")
    importlib.import_module("GPyOpt.experiment_design.base")
    rd = importlib.import_module("GPyOpt.experiment_design.random_design")

    def initial_design(design_name, space, init_points_count):   # experiment_design/__init__.py:8-24, 'random' branch
        if design_name != 'random':
            raise ValueError('Unknown design type: ' + design_name)
        return rd.RandomDesign(space).get_samples(init_points_count)
    sys.modules["GPyOpt.experiment_design"].initial_design = initial_design
    ns.initial_design = initial_design
    ns.space = importlib.import_module("GPyOpt.core.task.space")
    ns.objective = importlib.import_module("GPyOpt.core.task.objective")
    ns.cost = importlib.import_module("GPyOpt.core.task.cost")
    importlib.import_module("GPyOpt.util.duplicate_manager")
    ns.optimizer = importlib.import_module("GPyOpt.optimization.optimizer")
    ns.anchor = importlib.import_module("GPyOpt.optimization.anchor_points_generator")
    ns.acq_opt = importlib.import_module("GPyOpt.optimization.acquisition_optimizer")
    importlib.import_module("GPyOpt.core.evaluators.base")
    ns.sequential = importlib.import_module("GPyOpt.core.evaluators.sequential")
    sys.modules["GPyOpt.acquisitions"].AcquisitionLP = ns.AcquisitionLP
    ns.lp_eval = importlib.import_module("GPyOpt.core.evaluators.batch_local_penalization")
    # SciPy API drift (the reference pinned scipy 1.2): OptimizeResult.fun used to be the objective's raw (1, 1) array and
    # estimate_L indexes it as res.fun[0][0] (batch_local_penalization.py:67); modern SciPy returns a float.
    import scipy.optimize as _so

    def _minimize(*a, **kw):
        res = _so.minimize(*a, **kw)
        res.fun = np.atleast_2d(res.fun)
        return res
    ns.lp_eval.scipy = types.SimpleNamespace(optimize=types.SimpleNamespace(minimize=_minimize))
    ns.bo = importlib.import_module("GPyOpt.core.bo")
    return ns


def branin(X):
    """GPyOpt/GPyOpt/objective_examples/experiments2d.py:203-216 (sd = 0)."""
    X = np.atleast_2d(X)
    x1, x2 = X[:, 0], X[:, 1]
    b, c, r, s, t = 5.1 / (4 * np.pi ** 2), 5 / np.pi, 6, 10, 1 / (8 * np.pi)
    return ((x2 - b * x1 ** 2 + c * x1 - r) ** 2 + s * (1 - t) * np.cos(x1) + s).reshape(-1, 1)


BRANIN_DOMAIN = [{'name': 'x1', 'type': 'continuous', 'domain': (-5, 10)}, {'name': 'x2', 'type': 'continuous', 'domain': (1, 15)}]


def run_reference_bo(ns, f, domain, kernel_name, seed, iters, acquisition_type='EI', exact_feval=True, initial_design_numdata=5,
                     optimize_restarts=5, evaluator_type='sequential', batch_size=1):
    """methods/bayesian_optimization.py:76-171 wiring + BO.run_optimization (core/bo.py:73-168), all reference code."""
    np.random.seed(seed)
    space = ns.space.Design_space(domain, None)
    objective = ns.objective.SingleObjective(f, batch_size, 'no_name')
    cost = ns.cost.CostModel(None)
    X = ns.initial_design('random', space, initial_design_numdata)
    Y, _ = objective.evaluate(X)
    D = len(domain)
    kern = (ns.RBF if kernel_name == "rbf" else ns.Matern52)(D, variance=1.)
    model = ns.GPModel(kernel=kern, noise_var=None, exact_feval=exact_feval, optimizer='lbfgs', max_iters=1000,
                       optimize_restarts=optimize_restarts, sparse=False, num_inducing=10, verbose=False, ARD=False)
    aopt = ns.acq_opt.AcquisitionOptimizer(space, 'lbfgs', model=model)
    if acquisition_type == 'EI':
        acq = ns.AcquisitionEI(model, space, aopt, cost.cost_withGradients, 0.01)
    else:
        acq = ns.AcquisitionLCB(model, space, aopt, None, 2)
    if evaluator_type == 'local_penalization':
        evaluator = ns.lp_eval.LocalPenalization(ns.AcquisitionLP(model, space, aopt, acq, 'none'), batch_size)
    else:
        evaluator = ns.sequential.Sequential(acq)
    bo = ns.bo.BO(model=model, space=space, objective=objective, acquisition=evaluator.acquisition, evaluator=evaluator, X_init=X,
                  Y_init=Y, cost=cost, normalize_Y=True, model_update_interval=1, de_duplication=False)
    bo.run_optimization(max_iter=iters, verbosity=False)
    return bo


if __name__ == "__main__":
    ns = load_bo()
    bo = run_reference_bo(ns, branin, BRANIN_DOMAIN, "rbf", 0, 3)
    print(bo.X, bo.Y.ravel())
