import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from gaussian_process_optimization_b200 import GPy
import oracle_backend as OB
from test_hostapi import branin
np.random.seed(3)
xg1, xg2 = np.linspace(-5, 10, 5), np.linspace(0, 15, 5)
X = np.zeros((25, 2))
for i, x1 in enumerate(xg1):
    for j, x2 in enumerate(xg2):
        X[i + 5 * j, :] = [x1, x2]
Y = branin(X)
for backend in ("cuda", "oracle"):
    np.random.seed(3)
    k = GPy.kern.RBF(input_dim=2, ARD=True)
    m = GPy.models.GPRegression(X, Y, k) if backend == "cuda" else OB.oracle_gp_regression(X, Y, k)
    m.likelihood.variance.fix(1e-5)
    m.randomize()
    orig = m._objective_grads
    cnt = [0]
    def wrapped(x, orig=orig, m=m):
        try:
            r = orig(x)
        except Exception as e:
            print("EXC", x, repr(e)); raise
        cnt[0] += 1
        if cnt[0] < 60: print(backend, x, r[0], r[1])
        return r
    m._objective_grads = wrapped
    try:
        m.optimize()
    except Exception as e:
        print("optimize failed:", repr(e))
    print(m)
