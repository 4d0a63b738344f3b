"""NLL+grad with the products of the factorisation on the int8 tensor cores (experimental engine) vs the fp64 DMMA engine:
time per evaluation and agreement of the results, per threshold.  Writes gpurun_out/ozaki_fit_perf.json."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from bench import synth  # noqa: E402
from gaussian_process_optimization_b200 import native  # noqa: E402

SLICES = int(os.environ.get("OZAKI_SLICES", "7"))
sizes = [int(a) for a in sys.argv[1:]] or [4096, 8192, 16384]
out = {"slices": SLICES}
for N in sizes:
    D = 16
    X, Y, ls = synth(N, D)
    m = native.NativeModel("rbf", True, D, 1, n_cap=N, cand_block=1024)
    m.set_data(torch.from_numpy(X).cuda(), torch.from_numpy(Y).cuda())
    res = {}
    base = None
    mins = [int(t) for t in os.environ.get("OZAKI_MIN_LIST", "1024,2048,4096,8192,16384").split(",")]
    for min_n in [0] + [t for t in mins if t <= N]:
        native.set_ozaki(min_n, SLICES)
        ts = []
        for i in range(4):
            m.set_theta(1.0, ls, 1e-2)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            info, logL, g = m.fit(True)
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        if min_n == 0:
            base = (logL, g.copy())
        res["min_n_%d" % min_n] = {"ms": min(ts[1:]) * 1e3, "info": int(info), "logL_rel_vs_dmma": abs(logL - base[0]) / abs(base[0]),
                                   "grad_rel_vs_dmma": float(np.max(np.abs(g - base[1])) / np.max(np.abs(base[1])))}
        print(N, min_n, res["min_n_%d" % min_n], flush=True)
    native.set_ozaki(0, SLICES)
    out["N%d" % N] = res
    m.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/ozaki_fit_perf.json", "w"), indent=1)
