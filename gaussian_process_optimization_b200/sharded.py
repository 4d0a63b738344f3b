"""Candidate-set data parallelism for the acquisition step (north_star (c)).

The N x N factorisation stays on one GPU per rank (every rank holds the same model: same data, same hyper-parameters, so
the fit is replicated, not communicated); the candidate set is split into contiguous ranges (the `divide_data` idiom of
GPy/GPy/util/parallel.py:14-30), every rank scores its range with the fused device path and keeps its k best, and ONE small
all-gather of k * (value, global index, D coordinates) doubles per rank (NCCL over NVLink on GPUs, gloo in the CPU tests)
lets every rank form the same global top-k -- the anchor points of GPyOpt's AcquisitionOptimizer
(optimization/anchor_points_generator.py:58-63), ties resolved towards the lowest global index.
"""
import numpy as np


def divide_candidates(n_total, rank, world):
    """Contiguous [start, end) range of rank `rank` (first `n_total % world` ranks get one extra row)."""
    base, rem = divmod(int(n_total), int(world))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def merge_topk(vals, idx, pts, k):
    """Global k smallest by (value, index) from concatenated per-rank lists; deterministic."""
    vals = np.asarray(vals, dtype=np.float64).ravel()
    idx = np.asarray(idx, dtype=np.int64).ravel()
    pts = np.asarray(pts, dtype=np.float64).reshape(vals.size, -1)
    keep = idx >= 0
    vals, idx, pts = vals[keep], idx[keep], pts[keep]
    order = np.lexsort((idx, vals))[:k]
    return vals[order], idx[order], pts[order]


def all_gather_topk(vals, idx, pts, k, group=None, device=None):
    """All-gather the per-rank top-k and merge.  Works without an initialised process group (world size 1)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return merge_topk(vals, idx, pts, k)
    world = dist.get_world_size(group)
    d = np.asarray(pts).reshape(len(vals), -1).shape[1]
    kk = len(vals)
    # pad to k rows so that every rank contributes the same message size (a shard may hold fewer than k candidates)
    buf = np.full((k, 2 + d), np.nan)
    buf[:, 1] = -1.0
    buf[:kk, 0], buf[:kk, 1], buf[:kk, 2:] = vals, np.asarray(idx, dtype=np.float64), np.asarray(pts).reshape(kk, d)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    mine = torch.from_numpy(buf).to(device)
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine, group=group)
    allb = torch.stack(gathered).cpu().numpy().reshape(world * k, 2 + d)
    return merge_topk(allb[:, 0], allb[:, 1].astype(np.int64), allb[:, 2:], k)


def all_gather_topk_device(rows, k, group=None):
    """Device path of the same exchange: `rows` is the (k, d + 2) CUDA tensor NativeModel.acq_topk_dev left on the device
    ([f, global index, coordinates]; empty slots carry index -1).  The all-gather is issued from the device buffer behind the
    scoring kernels (same stream when the model runs on torch's current stream, ordered by NCCL's stream dependency otherwise) --
    no device -> host -> device round trip in front of the collective; ONE small copy of world * k rows to the host at the end."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        allb = rows.cpu().numpy()
    else:
        world = dist.get_world_size(group)
        if dist.get_backend(group) == "gloo":              # gloo has no CUDA all-gather (CPU tests, two ranks on one GPU): via the host
            mine = rows.cpu()
            parts = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(parts, mine, group=group)
            allb = torch.cat(parts).numpy()
        else:
            out = torch.empty((world * rows.shape[0], rows.shape[1]), dtype=rows.dtype, device=rows.device)
            dist.all_gather_into_tensor(out, rows.contiguous(), group=group)
            allb = out.cpu().numpy()
    return merge_topk(allb[:, 0], allb[:, 1].astype(np.int64), allb[:, 2:], k)


class ShardedAnchorScorer(object):
    """Score this rank's candidate shard and return the GLOBAL k best candidates (identical on every rank).

    `model` is a GPyOpt-level GPModel (or anything with `acquisition_topk(acq, par, X, k, index_offset)`); `score_fn` may
    replace the device path with any callable X -> scores (used by the CPU tests of the collective logic)."""

    def __init__(self, model=None, acq="EI", par=None, group=None, score_fn=None):
        self.model, self.acq, self.group, self.score_fn = model, acq, group, score_fn
        self.par = par if par is not None else (0.01 if acq == "EI" else 2.0)

    def local_topk(self, X_shard, k, index_offset):
        n = X_shard.shape[0]
        kk = min(k, n)
        if kk == 0:
            d = X_shard.shape[1]
            return np.zeros(0), np.zeros(0, dtype=np.int64), np.zeros((0, d))
        if self.score_fn is not None:
            s = np.asarray(self.score_fn(X_shard), dtype=np.float64).ravel()
            order = np.argsort(s, kind="stable")[:kk]
            return s[order], order.astype(np.int64) + index_offset, np.asarray(X_shard)[order]
        post = self.model.model.posterior if hasattr(self.model, "get_fmin") else self.model   # GPModel or a NativeModel
        fmin = self.model.get_fmin() if hasattr(self.model, "get_fmin") else post.fmin()
        return post.acq_topk(self.acq, self.par, fmin if self.acq == "EI" else 0.0, X_shard, kk, index_offset=index_offset)

    def topk(self, X_shard, k, index_offset):
        vals, idx, pts = self.local_topk(X_shard, k, index_offset)
        return all_gather_topk(vals, idx, pts, k, group=self.group)
