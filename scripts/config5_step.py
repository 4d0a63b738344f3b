"""BASELINE.json config 5 at its end state: batch Bayesian optimisation (local penalisation, batch = 64) on a synthetic 20-D
"Hartmann-style" objective (4 anisotropic Gaussian bumps, coefficients from RandomState(20)) with the model grown
to N = 32768, on all GPUs of one node:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        scripts/config5_step.py --n-end 32768 --batch 64 --iters 2

Every rank holds the same data, model and NumPy RNG state (the N x N factorisation is replicated, never communicated); the
ranks share the work that is independent given that state: one optimize_restarts restart per rank and one L-BFGS-B anchor
refinement per rank.  The second iteration's model update starts from an O(N^2 b) append of the 64 new points.
Not the bench contract: wall-clock phases of the BO loop, written to gpurun_out/config5_step.json by rank 0.
"""
import argparse
import hashlib
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from gaussian_process_optimization_b200 import GPy, GPyOpt, native  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n-end", type=int, default=32768)
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--iters", type=int, default=2)
ap.add_argument("--dim", type=int, default=20)
ap.add_argument("--max-iters", type=int, default=5, help="L-BFGS-B iterations per hyper-parameter restart")
ap.add_argument("--out", default="gpurun_out/config5_step.json")
args = ap.parse_args()

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

D = args.dim
rs = np.random.RandomState(20)
centres = rs.uniform(0.15, 0.85, (4, D))
widths = rs.uniform(0.35, 0.9, (4, D))
heights = np.array([1.0, 1.2, 3.0, 3.2])


def objective(X):
    X = np.atleast_2d(X)
    z = (X[:, None, :] - centres[None]) / widths[None]
    return -(heights[None] * np.exp(-0.5 * (z * z).sum(-1))).sum(-1, keepdims=True)


domain = [{'name': 'x%d' % q, 'type': 'continuous', 'domain': (0., 1.)} for q in range(D)]
n0 = args.n_end - args.batch * args.iters
X0 = np.random.RandomState(5).uniform(0, 1, (n0, D))
Y0 = objective(X0)
np.random.seed(0)

model = GPyOpt.models.GPModel(kernel=GPy.kern.Matern52(D, ARD=True), exact_feval=False, verbose=False, max_iters=args.max_iters,
                              optimize_restarts=max(world, 1), distributed_restarts=world > 1)
bo = GPyOpt.methods.BayesianOptimization(objective, domain=domain, model=model, X=X0, Y=Y0, acquisition_type='EI',
                                         evaluator_type='local_penalization', batch_size=args.batch, normalize_Y=True,
                                         distributed_anchors=world > 1)

phases = []


def timed(name, fn):
    def wrapper(*a, **k):
        torch.cuda.synchronize()
        c0, t0 = native.launch_count(), time.perf_counter()
        r = fn(*a, **k)
        torch.cuda.synchronize()
        phases.append({"phase": name, "n_train": int(bo.X.shape[0]), "seconds": time.perf_counter() - t0,
                       "kernel_launches": native.launch_count() - c0,
                       "appends": int(getattr(getattr(model.model, "inference_method", None), "n_appends", 0)) if model.model is not None else 0})
        if rank == 0:
            print(json.dumps(phases[-1]), flush=True)
        return r
    return wrapper


bo._update_model = timed("model_update (set_XY + optimize_restarts, one restart per rank)", bo._update_model)
bo._compute_next_evaluations = timed("batch of %d by local penalisation (one anchor refinement per rank)" % args.batch,
                                     bo._compute_next_evaluations)
t0 = time.perf_counter()
bo.run_optimization(max_iter=args.iters)
total = time.perf_counter() - t0

digest = hashlib.sha256(np.ascontiguousarray(bo.X).tobytes()).hexdigest()
same = True
if world > 1:
    mine = torch.tensor([int(digest[:15], 16)], dtype=torch.int64, device="cuda")
    allv = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allv, mine)
    same = len({int(v.item()) for v in allv}) == 1
if rank == 0:
    Xb = bo.X[n0:]
    dmin = min(np.sqrt(((Xb[i] - Xb[j]) ** 2).sum()) for i in range(len(Xb)) for j in range(i))
    out = {"config": "BASELINE config 5: batch BO, local penalisation, batch=%d, D=%d, N grown %d -> %d" % (args.batch, D, n0, bo.X.shape[0]),
           "n_gpus": world, "total_seconds": total, "phases": phases, "trajectory_identical_on_all_ranks": bool(same),
           "best_objective_initial": float(Y0.min()), "best_objective_final": float(bo.Y.min()),
           "min_distance_within_proposed_points": float(dmin), "hyper_parameters": [float(v) for v in model.model[:]],
           "appends_used": int(model.model.inference_method.n_appends), "max_iters_per_restart": args.max_iters}
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(out, open(args.out, "w"), indent=1)
    print(json.dumps({k: v for k, v in out.items() if k != "phases"}), flush=True)
assert same, "ranks disagree on the BO trajectory"
if world > 1:
    dist.destroy_process_group()
