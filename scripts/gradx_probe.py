"""Times the EI value + gradient pass over 6 candidate blocks (2048 each) against the N = 16384 model, with the DMMA engine and
with the int8 engine, and prints a checksum of the acquisition gradients so that kernel variants (GPB_GX_VARIANT) can be compared."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from bench import synth  # noqa: E402
from gaussian_process_optimization_b200 import native  # noqa: E402

N, D = 16384, int(os.environ.get("GPB_PROBE_D", "16"))
X, Y, ls = synth(N, D)
out = {"D": D, "tile": os.environ.get("GPB_GRADX_TILE", "1")}
Xc = np.random.RandomState(4321).uniform(0, 1, (2048 * 6, D))
Xd = torch.from_numpy(Xc).cuda()
for engine in os.environ.get("GPB_PROBE_ENGINES", "dmma,int8").split(","):
    native.set_ozaki(8192 if engine == "int8" else 0, 16)
    m = native.NativeModel("mat52", True, D, 1, n_cap=N, cand_block=2048)
    m.set_data(X, Y)
    m.set_theta(1.0, ls, 1e-6)
    assert m.fit(False)[0] == 0
    fmin = m.fmin()
    m.acq_topk_dev("EI", 0.01, fmin, Xd[:4096], 5, with_gradients=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = m.acq_topk_dev("EI", 0.01, fmin, Xd, 5, with_gradients=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    a = m.acquisition("EI", 0.01, fmin, Xc[:512], with_gradients=True, want_moments=True)
    out[engine] = {"ms_6_blocks": ms, "cand_per_s": 2048 * 6 / ms * 1e3,
                   "df_sum": float(np.abs(a["df"]).sum()), "dmdx_sum": float(np.abs(a["dmdx"]).sum()),
                   "dsdx_sum": float(np.abs(a["dsdx"]).sum())}
    m.close()
print(json.dumps(out), flush=True)
