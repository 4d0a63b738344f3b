"""Short program for ncu: two 4096^3 products through the int8 engine in its modular mode (16 moduli), nothing else.
  ncu --set full --clock-control none --import-source on -k regex:oz_crt_combine_kernel --launch-skip 1 -c 1 python scripts/ozaki_crt_ncu_probe.py"""
import sys

import torch

sys.path.insert(0, ".")
from gaussian_process_optimization_b200 import native  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
A = torch.randn(n, n, dtype=torch.float64, device="cuda")
B = torch.randn(n, n, dtype=torch.float64, device="cuda")
C = torch.empty(n, n, dtype=torch.float64, device="cuda")
for _ in range(2):
    native.ozaki_dgemm(0, 0, 1.0, A, B, 0.0, C, slices=16)
torch.cuda.synchronize()
print("ok", float(C[0, 0]))
