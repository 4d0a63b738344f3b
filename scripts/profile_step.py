"""Short program for ncu: two NLL+grad evaluations at the headline size (first one warms up), nothing else."""
import sys

import numpy as np

sys.path.insert(0, ".")
from bench import synth, N_TRAIN, DIM, KIND  # noqa: E402
from gaussian_process_optimization_b200 import native  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else N_TRAIN
X, Y, ls = synth(N, DIM)
m = native.NativeModel(KIND, True, DIM, 1, n_cap=N, cand_block=128)
m.set_data(X, Y)
m.set_theta(1.0, ls, 1e-2)
for i in range(2):
    c0 = native.launch_count()
    info, logL, g = m.fit(True)
    print("fit", i, "info", info, "logL", logL, "launches", native.launch_count() - c0, flush=True)
m.close()
