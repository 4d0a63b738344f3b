"""BASELINE.json config 4 at full size: EI value + gradient over 2^24 synthetic candidates (D = 16) against the N = 16384
Matern52 model with the exact_feval noise, sharded over the ranks of one node; per-shard top-5 all-gathered.

    python -m torch.distributed.run --nproc-per-node G scripts/ei_2p24.py [chunks] [EI|LCB]   (chunks of 2^20 rows, default 16)

Chunk c = RandomState(4321 + c).uniform(0, 1, (2^20, 16)); rank r takes the contiguous chunk range r*C/G .. (r+1)*C/G.
Every rank holds the same fitted model (the fit is replicated); values and gradients stay in HBM, only the top-5 travel."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from bench import synth  # noqa: E402
from gaussian_process_optimization_b200 import native, sharded  # noqa: E402

N, D, CHUNK = 16384, 16, 2 ** 20
chunks = int(sys.argv[1]) if len(sys.argv) > 1 else 16
ACQ = sys.argv[2] if len(sys.argv) > 2 else "EI"                                   # "EI" (jitter 0.01) or "LCB" (weight 2)
PAR = 0.01 if ACQ == "EI" else 2.0
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
X, Y, ls = synth(N, D)
m = native.NativeModel("mat52", True, D, 1, n_cap=N, cand_block=4096)
m.set_data(X, Y)
m.set_theta(1.0, ls, 1e-6)
info, logL, _ = m.fit(False)
assert info == 0
fmin = m.fmin()
lo, hi = sharded.divide_candidates(chunks, rank, world)
host = [np.random.RandomState(4321 + c).uniform(0, 1, (CHUNK, D)) for c in range(lo, hi)]
m.acq_topk_full(ACQ, PAR, fmin, torch.from_numpy(host[0][:8192]).cuda(), 5)      # warm-up
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
best = []
keep = []
for i, c in enumerate(range(lo, hi)):
    Xc = torch.from_numpy(host[i]).cuda()                                            # H2D of the chunk inside the timed region
    vals, idx, pts, f, df = m.acq_topk_full(ACQ, PAR, fmin, Xc, 5, index_offset=c * CHUNK)
    best.append((vals, idx, pts))
    keep.append((f, df))                                                              # results stay resident
vals = np.concatenate([b[0] for b in best]) if best else np.zeros(0)
idx = np.concatenate([b[1] for b in best]) if best else np.zeros(0, dtype=np.int64)
pts = np.concatenate([b[2] for b in best]) if best else np.zeros((0, D))
v5, i5, p5 = sharded.merge_topk(vals, idx, pts, 5)
g5 = sharded.all_gather_topk(v5, i5, p5, 5)
e1.record()
torch.cuda.synchronize()
t = e0.elapsed_time(e1) * 1e-3
if world > 1:
    tt = torch.tensor([t], dtype=torch.float64, device="cuda")
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t = float(tt[0])
if rank == 0:
    ncand = chunks * CHUNK
    fl = 2.0 * N ** 2 + N * (6 * D + 40)
    res = {"acquisition": ACQ, "candidates": ncand, "gpus": world, "seconds": t, "candidates_per_s": ncand / t, "algorithmic_tflops_total": fl * ncand / t / 1e12,
           "algorithmic_tflops_per_gpu": fl * ncand / t / 1e12 / world, "top5_idx": [int(i) for i in g5[1]], "top5_f": [float(v) for v in g5[0]],
           "fmin": fmin, "logL": logL, "wall_s": time.perf_counter() - t0}
    print(json.dumps(res), flush=True)
    json.dump(res, open("gpurun_out/%s_2p24_g%d_c%d.json" % (ACQ.lower(), world, chunks), "w"), indent=1)
m.close()
if world > 1:
    dist.destroy_process_group()
