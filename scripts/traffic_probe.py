"""One NLL+grad evaluation at the headline size, for `ncu -k regex:"gemm_dmma_kernel<1, 1"` (the Ky^-1 = M^T M launch, the dominant
one): DRAM bytes and duration per value of GPB_TRI_BAND (block rows per band of the lower-tile enumeration)."""
import sys

import torch

sys.path.insert(0, ".")
from bench import synth, N_TRAIN, DIM  # noqa: E402
from gaussian_process_optimization_b200 import native  # noqa: E402

X, Y, ls = synth(N_TRAIN, DIM)
m = native.NativeModel("rbf", True, DIM, 1, n_cap=N_TRAIN, cand_block=128)
m.set_data(torch.from_numpy(X).cuda(), torch.from_numpy(Y).cuda())
m.set_theta(1.0, ls, 1e-2)
info, logL, g = m.fit(True)
assert info == 0
print("logL", logL)
m.close()
