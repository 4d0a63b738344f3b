"""Sweep of the two-stream factorisation threshold (gpb_set_overlap): NLL+grad time and bitwise equality of the results."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from bench import synth  # noqa: E402
from gaussian_process_optimization_b200 import native  # noqa: E402

for kind, N, D in (("rbf", 4096, 8), ("rbf", 16384, 16)):
    X, Y, ls = synth(N, D)
    m = native.NativeModel(kind, True, D, 1, n_cap=N, cand_block=128)
    m.set_data(torch.from_numpy(X).cuda(), torch.from_numpy(Y).cuda())
    ref = None
    for ov in (0, 256, 512, 1024, 2048, 4096):
        native.set_overlap(ov)
        ts = []
        for i in range(4):
            m.set_theta(1.0, ls, 1e-2)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            info, logL, g = m.fit(True)
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        if ref is None:
            ref = (logL, g.copy())
        same = (logL == ref[0]) and np.array_equal(g, ref[1])
        print("N=%d overlap_min_n=%d: %.3f ms (min of 3), bitwise equal to single-stream: %s" % (N, ov, min(ts[1:]) * 1e3, same), flush=True)
    m.close()
