"""Short program for ncu: one fitted model at the headline size, then a few M = 1 (and M = 5) EI value+gradient calls -- the calls
L-BFGS-B makes from every anchor point (optimizer.py:46-51)."""
import sys

import numpy as np

sys.path.insert(0, ".")
from bench import synth, DIM  # noqa: E402
from gaussian_process_optimization_b200 import native  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
X, Y, ls = synth(N, DIM)
m = native.NativeModel("mat52", True, DIM, 1, n_cap=N, cand_block=128)
m.set_data(X, Y)
m.set_theta(1.0, ls, 1e-2)
assert m.fit(False)[0] == 0
fmin = m.fmin()
rs = np.random.RandomState(3)
for mc in (1, 1, 1, 5):
    c0 = native.launch_count()
    r = m.acquisition("EI", 0.01, fmin, rs.uniform(0, 1, (mc, DIM)), with_gradients=True)
    print("M =", mc, "launches", native.launch_count() - c0, "f", r["f"].ravel()[:2], flush=True)
m.close()
