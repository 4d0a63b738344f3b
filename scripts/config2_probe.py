"""BASELINE config 2 (RBF-ARD, N = 4096, D = 8): NLL + gradient time per evaluation (device-resident inputs, best of 20 after warm-up),
for the knob values given in the environment; also cuSOLVER's potrf alone through torch as a yardstick.  One JSON line."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from bench import synth  # noqa: E402
from gaussian_process_optimization_b200 import native  # noqa: E402

N, D = int(os.environ.get("PROBE_N", "4096")), 8
X, Y, ls = synth(N, D)
m = native.NativeModel("rbf", True, D, 1, n_cap=N, cand_block=128)
m.set_data(torch.from_numpy(X).cuda(), torch.from_numpy(Y).cuda())
ts = []
for i in range(26):
    m.set_theta(1.0 + 1e-3 * (i % 5), ls, 1e-2)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    info, logL, g = m.fit(True)
    b.record()
    torch.cuda.synchronize()
    assert info == 0
    if i >= 6:
        ts.append(a.elapsed_time(b))
out = {"N": N, "ms_best": min(ts), "ms_median": float(np.median(ts)), "launches_per_eval": None,
       "env": {k: v for k, v in os.environ.items() if k.startswith("GPB_")}}
c0 = native.launch_count()
m.set_theta(1.0, ls, 1e-2)
m.fit(True)
out["launches_per_eval"] = native.launch_count() - c0
if os.environ.get("PROBE_CUSOLVER"):
    K = torch.from_numpy(np.asarray(native.kern_K("rbf", X, None, 1.0, ls))).cuda() + 1e-2 * torch.eye(N, dtype=torch.float64, device="cuda")
    tt = []
    for i in range(8):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        L = torch.linalg.cholesky(K)
        b.record()
        torch.cuda.synchronize()
        tt.append(a.elapsed_time(b))
    out["cusolver_potrf_only_ms"] = min(tt[2:])
    tt = []
    for i in range(6):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        Ki = torch.cholesky_inverse(L)
        b.record()
        torch.cuda.synchronize()
        tt.append(a.elapsed_time(b))
    out["cusolver_potri_only_ms"] = min(tt[2:])
print(json.dumps(out), flush=True)
