"""Short program for ncu: one fitted model at the headline size, then a few M = 1 / 5 / 8 EI value+gradient calls -- the calls
L-BFGS-B makes from every anchor point (optimizer.py:46-51), coalesced by the host's LockstepEvaluator.
Prints wall time per call for the fused cooperative kernel and (GPB_SKINNY_FUSED=0) the multi-kernel route."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from bench import synth, DIM  # noqa: E402
from gaussian_process_optimization_b200 import native  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
X, Y, ls = synth(N, DIM)
m = native.NativeModel("mat52", True, DIM, 1, n_cap=N, cand_block=128)
m.set_data(X, Y)
m.set_theta(1.0, ls, 1e-2)
assert m.fit(False)[0] == 0
fmin = m.fmin()
rs = np.random.RandomState(3)
for mc in (1, 1, 5, 8):
    xs = rs.uniform(0, 1, (reps + 2, mc, DIM))
    c0 = native.launch_count()
    r = m.acquisition("EI", 0.01, fmin, xs[0], with_gradients=True)
    nl = native.launch_count() - c0
    m.acquisition("EI", 0.01, fmin, xs[1], with_gradients=True)
    t0 = time.perf_counter()
    for i in range(reps):
        m.acquisition("EI", 0.01, fmin, xs[2 + i], with_gradients=True)
    t = (time.perf_counter() - t0) / reps
    print("M =", mc, "launches", nl, "ms/call %.4f" % (t * 1e3), "f", r["f"].ravel()[:2], flush=True)
m.close()
