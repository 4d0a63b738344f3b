"""BASELINE config 1 (GPyOpt BayesianOptimization EI on 2-D Branin, GPRegression RBF, 30 iterations, seed 0): wall time of the whole
BO loop through the host mirror on the CUDA path, and -- same host code, same seed -- on the CPU oracle backend of the tests
(tests/oracle_backend.py; the reference's own arithmetic restated in NumPy/SciPy).  Small-N regime: launch latency and host
overhead, not throughput."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from gaussian_process_optimization_b200 import GPy, GPyOpt  # noqa: E402
import oracle_backend as OB  # noqa: E402


def branin(X):
    X = np.atleast_2d(X)
    x1, x2 = X[:, 0], X[:, 1]
    b, c, r, s, t = 5.1 / (4 * np.pi ** 2), 5 / np.pi, 6, 10, 1 / (8 * np.pi)
    return ((x2 - b * x1 ** 2 + c * x1 - r) ** 2 + s * (1 - t) * np.cos(x1) + s).reshape(-1, 1)


DOMAIN = [{'name': 'x1', 'type': 'continuous', 'domain': (-5, 10)}, {'name': 'x2', 'type': 'continuous', 'domain': (1, 15)}]
out = {}
for backend in ("cuda", "cuda", "oracle"):           # the first CUDA run warms up (context, module load)
    np.random.seed(0)
    kw = dict(kernel=GPy.kern.RBF(2), exact_feval=True, verbose=False)
    model = GPyOpt.models.GPModel(**kw) if backend == "cuda" else OB.OracleGPModel(**kw)
    t0 = time.perf_counter()
    bo = GPyOpt.methods.BayesianOptimization(branin, domain=DOMAIN, model=model, acquisition_type='EI', exact_feval=True,
                                             initial_design_numdata=5, initial_design_type='random')
    bo.run_optimization(max_iter=30)
    dt = time.perf_counter() - t0
    from gaussian_process_optimization_b200 import native
    out[backend] = {"seconds": dt, "evaluations": int(bo.X.shape[0]), "best": float(bo.Y.min()),
                    "kernel_launches_total": int(native.launch_count()) if backend == "cuda" else 0,
                    "hyperparameter_evaluations": int(sum(r.funct_eval for r in model.model.optimization_runs))}
    print(backend, out[backend], flush=True)
out["same_trajectory"] = bool(abs(out["cuda"]["best"] - out["oracle"]["best"]) < 1e-6)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/config1_time.json", "w"), indent=1)
