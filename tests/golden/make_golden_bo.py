"""Generate tests/golden/bo/*.npz by running the reference's OWN BO loop (ref_bo_harness.py).  Container-only.

Cases: BASELINE.json config 1 (BO on 2-D Branin, GPRegression RBF, EI, exact_feval, 5 random initial points, seed 0, 30
iterations), the same with LCB (15 iterations), and a local-penalisation batch run (Matern52, batch 3, 3 iterations)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_bo_harness as rb  # noqa: E402

CASES = [
    # name, kernel, acquisition, seed, iterations, initial points, restarts, evaluator, batch
    ("branin_rbf_ei_seed0", "rbf", "EI", 0, 30, 5, 5, "sequential", 1),
    ("branin_rbf_lcb_seed1", "rbf", "LCB", 1, 15, 5, 5, "sequential", 1),
    ("branin_mat52_ei_lp_batch3_seed4", "mat52", "EI", 4, 3, 6, 1, "local_penalization", 3),
]


def main():
    ns = rb.load_bo()
    for name, kern, acq, seed, iters, n0, restarts, ev, batch in CASES:
        bo = rb.run_reference_bo(ns, rb.branin, rb.BRANIN_DOMAIN, kern, seed, iters, acquisition_type=acq, exact_feval=True,
                                 initial_design_numdata=n0, optimize_restarts=restarts, evaluator_type=ev, batch_size=batch)
        np.savez_compressed(os.path.join(HERE, "bo", name + ".npz"), X=bo.X, Y=bo.Y, theta=bo.model_parameters_iterations,
                            kernel=kern, acquisition=acq, seed=seed, iters=iters, n0=n0, restarts=restarts, evaluator=ev, batch=batch)
        print("%-36s evaluations %d  best %.6f at %s" % (name, bo.X.shape[0], bo.Y.min(), bo.X[np.argmin(bo.Y)]))


if __name__ == "__main__":
    main()
