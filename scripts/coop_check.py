import os, sys, json
import numpy as np, torch
sys.path.insert(0, ".")
from bench import synth
from gaussian_process_optimization_b200 import native
N = int(sys.argv[1]); D = 8 if N <= 4096 else 16
X, Y, ls = synth(N, D)
m = native.NativeModel("rbf", True, D, 1, n_cap=N, cand_block=128)
m.set_data(torch.from_numpy(X).cuda(), torch.from_numpy(Y).cuda())
ts = []
for i in range(16):
    m.set_theta(1.0 + 1e-3 * (i % 5), ls, 1e-2)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); info, logL, g = m.fit(True); b.record(); torch.cuda.synchronize()
    assert info == 0
    if i >= 4: ts.append(a.elapsed_time(b))
m.set_theta(1.0, ls, 1e-2)
info, logL, g = m.fit(True)
al = m.get("alpha")
print(json.dumps({"N": N, "coop": os.environ.get("GPB_COOP_N", "default"), "graph": os.environ.get("GPB_GRAPH_MAX_NP", "default"), "ms_best": min(ts), "logL": repr(logL), "gsum": repr(float(np.sum(g))), "asum": repr(float(np.sum(al)))}))
