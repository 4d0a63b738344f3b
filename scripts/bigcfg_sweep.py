"""NLL+grad time for each large-launch GEMM configuration (gpb_gemm_config(100 + cfg))."""
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
    for cfg in [int(c) for c in sys.argv[1].split(",")]:
        native.gemm_config(100 + cfg)
        ts = []
        for i in range(4):
            m.set_theta(1.0, ls, 1e-2)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            info, logL, g = m.fit(True)
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        print("N=%d big cfg %d: %.3f ms  logL %.12g" % (N, cfg, min(ts[1:]) * 1e3, logL), flush=True)
    m.close()
