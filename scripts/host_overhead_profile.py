"""cProfile of the host side of one NLL+grad evaluation through the mirror classes (small N: the device part is ~0.1 ms)."""
import cProfile
import pstats
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from bench import synth  # noqa: E402
from gaussian_process_optimization_b200 import GPy  # noqa: E402

N, D = (int(sys.argv[1]) if len(sys.argv) > 1 else 128), 8
X, Y, ls = synth(N, D)
m = GPy.models.GPRegression(X, Y, kernel=GPy.kern.Matern52(D, ARD=True), noise_var=0.05)
x0 = m.optimizer_array.copy()
xs = [x0 + 1e-3 * np.random.RandomState(i).randn(x0.size) for i in range(2000)]
for x in xs[:50]:
    m._objective_grads(x)
t0 = time.perf_counter()
for x in xs:
    m._objective_grads(x)
print("N=%d: %.1f us per _objective_grads call" % (N, (time.perf_counter() - t0) / len(xs) * 1e6))
pr = cProfile.Profile()
pr.enable()
for x in xs:
    m._objective_grads(x)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
