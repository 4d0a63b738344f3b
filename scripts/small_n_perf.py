"""NLL+grad wall time at small N (launch-latency regime) and the kernel count per evaluation."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from bench import synth  # noqa: E402
from gaussian_process_optimization_b200 import native  # noqa: E402

for N in (32, 128, 256, 512, 1024, 2048, 4096, 8192):
    D = 8
    X, Y, ls = synth(N, D)
    m = native.NativeModel("rbf", True, D, 1, n_cap=N, cand_block=128)
    m.set_data(torch.from_numpy(X).cuda(), torch.from_numpy(Y).cuda())
    ts = []
    for i in range(12):
        m.set_theta(1.0 + 1e-3 * i, ls, 1e-2)
        c0 = native.launch_count()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        info, logL, g = m.fit(True)
        ts.append(time.perf_counter() - t0)
        nl = native.launch_count() - c0
    print("N=%5d: %.3f ms per NLL+grad (median of 12), %d kernel launches" % (N, np.median(ts) * 1e3, nl), flush=True)
    m.close()
