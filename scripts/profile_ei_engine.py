"""Short program for ncu: EI value + gradient over a few candidate blocks with the int8 engine in modular mode (18 moduli in the
predictive products; argv[1] = 0 for the fp64 DMMA path) against the N = 16384 model -- the launch list shows where a block's time goes.
  ncu --nvtx --nvtx-include "gpb_model_acq_topk_dev/" --metrics gpu__time_duration.sum --clock-control none --csv ... (both passes)"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from bench import synth  # noqa: E402
from gaussian_process_optimization_b200 import native  # noqa: E402

N, D = 16384, 16
X, Y, ls = synth(N, D)
native.set_ozaki(int(sys.argv[1]) if len(sys.argv) > 1 else 8192, 16)      # 0: the fp64 DMMA path
m = native.NativeModel("mat52", True, D, 1, n_cap=N, cand_block=2048)
m.set_data(X, Y)
m.set_theta(1.0, ls, 1e-6)
assert m.fit(False)[0] == 0
fmin = m.fmin()
Xd = torch.from_numpy(np.random.RandomState(4321).uniform(0, 1, (2048 * 6, D))).cuda()
m.acq_topk_dev("EI", 0.01, fmin, Xd[:4096], 5, with_gradients=True)
torch.cuda.synchronize()
c0 = native.launch_count()
m.acq_topk_dev("EI", 0.01, fmin, Xd, 5, with_gradients=True)
torch.cuda.synchronize()
print("launches of the 6-block pass:", native.launch_count() - c0, flush=True)
m.close()
