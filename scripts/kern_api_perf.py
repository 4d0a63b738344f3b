"""Latency of the stateless Kern-contract entry points (host arrays in / out) at small sizes."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from gaussian_process_optimization_b200 import GPy  # noqa: E402

rs = np.random.RandomState(0)
for n, d in ((50, 2), (500, 8), (2000, 8)):
    X = rs.uniform(0, 1, (n, d))
    k = GPy.kern.Matern52(d, ARD=True)
    G = rs.randn(n, n)
    for name, fn in (("K", lambda: k.K(X)), ("update_gradients_full", lambda: k.update_gradients_full(G, X)),
                     ("gradients_X", lambda: k.gradients_X(G, X))):
        for _ in range(3):
            fn()
        t0 = time.perf_counter()
        for _ in range(20):
            fn()
        print("n=%d d=%d %s: %.3f ms per call" % (n, d, name, (time.perf_counter() - t0) / 20 * 1e3), flush=True)
