"""Quick device timings (not the bench contract): NLL+grad at the named sizes, and acquisition latency for small candidate counts."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from bench import synth  # noqa: E402
from gaussian_process_optimization_b200 import native  # noqa: E402

out = {}
sizes = [("rbf", 4096, 8), ("rbf", 16384, 16)] if len(sys.argv) < 2 else [("rbf", int(sys.argv[1]), 16)]
for kind, N, D in sizes:
    X, Y, ls = synth(N, D)
    m = native.NativeModel(kind, True, D, 1, n_cap=N, cand_block=1024)
    m.set_data(torch.from_numpy(X).cuda(), torch.from_numpy(Y).cuda())
    ts = []
    for i in range(5):
        m.set_theta(1.0 + 1e-3 * i, ls, 1e-2)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        info, logL, g = m.fit(True)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    flops = N ** 3 + N ** 2 * (6 * D + 62)
    out["fit_%s_%d_%d" % (kind, N, D)] = {"ms": min(ts) * 1e3, "tflops": flops / min(ts) / 1e12, "logL": logL}
    print(kind, N, D, ["%.3f" % (t * 1e3) for t in ts], logL, flush=True)
    fmin = m.fmin()
    rs = np.random.RandomState(5)
    for mc in (1, 4, 8, 9, 128, 1024):
        Xc = rs.uniform(0, 1, (mc, D))
        tt = []
        for i in range(4):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            m.acquisition("EI", 0.01, fmin, Xc, with_gradients=True)
            torch.cuda.synchronize()
            tt.append(time.perf_counter() - t0)
        out["acq_grad_N%d_mc%d_ms" % (N, mc)] = min(tt) * 1e3
        print("  acq+grad mc=%d: %.3f ms" % (mc, min(tt) * 1e3), flush=True)
    m.close()
json.dump(out, open("gpurun_out/fit_perf.json", "w"), indent=1)
