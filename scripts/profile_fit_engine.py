"""Short program for ncu: one NLL + gradient evaluation at N = 16384 with the products of the factorisation on the int8 engine
(16 moduli, levels >= 8192) -- the launch list shows what the evaluation spends outside the int8 products."""
import sys

import torch

sys.path.insert(0, ".")
from bench import synth  # noqa: E402
from gaussian_process_optimization_b200 import native  # noqa: E402

N, D = 16384, 16
X, Y, ls = synth(N, D)
native.set_ozaki(int(sys.argv[1]) if len(sys.argv) > 1 else 8192, 16)
m = native.NativeModel("mat52", True, D, 1, n_cap=N, cand_block=1024)
m.set_data(X, Y)
for i in range(2):
    m.set_theta(1.0, ls, 1e-2)
    torch.cuda.synchronize()
    c0 = native.launch_count()
    assert m.fit(True)[0] == 0
    torch.cuda.synchronize()
print("launches of one evaluation:", native.launch_count() - c0, flush=True)
m.close()
