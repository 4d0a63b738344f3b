"""The drop-in boundary seen from the reference's side (build container only: needs /root/reference).

The reference's OWN BO loop (GPyOpt core/bo.py, acquisition optimiser, EI, Sequential evaluator -- imported unmodified through
tests/golden/ref_bo_harness.py) is handed this repo's GPModel as its `model` (the BOModel plug-in point of
GPyOpt/GPyOpt/models/base.py:7-33 and methods/bayesian_optimization.py:125-133).  On CPU the GPModel runs on the oracle backend;
the run must follow the all-reference run evaluation by evaluation, i.e. the host mirror honours the contract the reference's
callers rely on: updateModel / predict / predict_withGradients / get_fmin / get_model_parameters / model.model.X, Y.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
pytestmark = pytest.mark.reference

SCRIPT = r'''
import sys, os
import numpy as np
sys.path.insert(0, %(root)r); sys.path.insert(0, %(here)r); sys.path.insert(0, os.path.join(%(here)r, "golden"))
import ref_bo_harness as rb
ns = rb.load_bo()
from gaussian_process_optimization_b200 import GPy
import oracle_backend as OB

def run(inject):
    np.random.seed(3)
    space = ns.space.Design_space(rb.BRANIN_DOMAIN, None)
    objective = ns.objective.SingleObjective(rb.branin, 1, 'no_name')
    cost = ns.cost.CostModel(None)
    X = ns.initial_design('random', space, 5)
    Y, _ = objective.evaluate(X)
    if inject:
        model = OB.OracleGPModel(kernel=GPy.kern.Matern52(2, variance=1.), exact_feval=True, verbose=False, optimize_restarts=2)
    else:
        model = ns.GPModel(kernel=ns.Matern52(2, variance=1.), exact_feval=True, optimize_restarts=2, verbose=False)
    aopt = ns.acq_opt.AcquisitionOptimizer(space, 'lbfgs', model=model)
    acq = ns.AcquisitionEI(model, space, aopt, cost.cost_withGradients, 0.01)         # the reference's EI on either model
    bo = ns.bo.BO(model=model, space=space, objective=objective, acquisition=acq, evaluator=ns.sequential.Sequential(acq),
                  X_init=X, Y_init=Y, cost=cost, normalize_Y=True)
    bo.run_optimization(max_iter=6)
    return bo.X, bo.Y, bo.model_parameters_iterations

a = run(False)
b = run(True)
np.savez(%(out)r, Xr=a[0], Yr=a[1], Xi=b[0], Yi=b[1])
'''


@pytest.mark.skipif(not os.path.isdir("/root/reference/GPyOpt/GPyOpt"), reason="needs the reference checkout")
def test_reference_bo_loop_accepts_the_b200_gpmodel(tmp_path):
    out = str(tmp_path / "inj.npz")
    # separate interpreter: the harness registers skeleton GPy / GPyOpt packages in sys.modules
    code = SCRIPT % {"root": os.path.dirname(HERE), "here": HERE, "out": out}
    subprocess.run([sys.executable, "-c", code], check=True, timeout=900)
    z = np.load(out)
    assert z["Xi"].shape == z["Xr"].shape == (11, 2)
    np.testing.assert_allclose(z["Xi"], z["Xr"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(z["Yi"], z["Yr"], rtol=1e-5, atol=1e-6)
