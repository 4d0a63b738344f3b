"""BASELINE.json config 3: GPRegression Matern52-ARD N=16384 D=16, full hyper-parameter optimize (L-BFGS-B on the host, every
objective/gradient evaluation one device round trip).  Reports evaluations, wall time and the log-likelihood trace."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from bench import synth  # noqa: E402
from gaussian_process_optimization_b200 import GPy  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
D = 16
max_iters = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
X, Y, ls = synth(N, D)
k = GPy.kern.Matern52(D, variance=1.0, lengthscale=ls, ARD=True)
t0 = time.perf_counter()
m = GPy.models.GPRegression(X, Y, kernel=k, noise_var=1e-2)
t_build = time.perf_counter() - t0
trace = []
orig = m._objective_grads


def traced(x):
    t = time.perf_counter()
    r = orig(x)
    trace.append((time.perf_counter() - t, float(r[0])))
    return r


m._objective_grads = traced
t0 = time.perf_counter()
run = m.optimize(optimizer='lbfgs', max_iters=max_iters)
t_opt = time.perf_counter() - t0
res = {"N": N, "D": D, "kernel": "Matern52-ARD", "construct_s": t_build, "optimize_s": t_opt, "evaluations": len(trace),
       "s_per_evaluation": float(np.mean([t for t, _ in trace])), "objective_first": trace[0][1], "objective_last": trace[-1][1],
       "status": run.status, "variance": float(k.variance.values[0]), "lengthscale": k.lengthscale.values.tolist(),
       "noise": float(m.likelihood.variance.values[0])}
print(json.dumps(res))
json.dump(res, open("gpurun_out/optimize_config3_N%d.json" % N, "w"), indent=1)
