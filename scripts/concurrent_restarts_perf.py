"""optimize_restarts on one GPU: sequential loop (the reference's) against K restarts at a time in host threads (one model copy
and CUDA stream each).  Wall time of 5 restarts x max_iters L-BFGS-B iterations at several training-set sizes; same runs."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from bench import synth  # noqa: E402
from gaussian_process_optimization_b200 import GPy  # noqa: E402

out = {}
D, restarts, iters = 8, 5, 15
for N in (128, 512, 1024, 2048, 4096):
    X, Y, ls = synth(N, D)
    row = {}
    for conc in (0, 5, 0, 5):
        m = GPy.models.GPRegression(X, Y, kernel=GPy.kern.Matern52(D, ARD=True), noise_var=0.05)
        m.Gaussian_noise.constrain_bounded(1e-9, 1e6, warning=False)
        np.random.seed(0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        m.optimize_restarts(num_restarts=restarts, optimizer='lbfgs', max_iters=iters, verbose=False, concurrent=conc)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        evals = sum(r.funct_eval for r in m.optimization_runs)
        key = "concurrent" if conc else "sequential"
        row[key] = {"seconds": min(dt, row.get(key, {}).get("seconds", 1e30)), "evaluations": int(evals),
                    "best_f": float(min(r.f_opt for r in m.optimization_runs))}
    row["speedup"] = row["sequential"]["seconds"] / row["concurrent"]["seconds"]
    row["same_optimum"] = bool(row["sequential"]["best_f"] == row["concurrent"]["best_f"])
    out["N%d" % N] = row
    print(N, json.dumps(row), flush=True)
json.dump(out, open("gpurun_out/concurrent_restarts_perf.json", "w"), indent=1)
