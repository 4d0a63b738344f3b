"""Round-2 additions on the device path: posterior generation stamps, empty top-k slots, the input-dimension cap, and the
multi-rank state distribution (rank 0 fits, theta / alpha / L^-1 are broadcast, the other ranks adopt them instead of refitting;
SURVEY.md 8e) with the device-side all-gather of the per-shard anchors.

The two-rank test gives every rank its own GPU over NCCL when the box has two; on a single-GPU box (the driver's test box) it runs both
ranks on ONE GPU over the gloo backend (NCCL refuses two ranks on one device) -- the same code path as `bench.py --gpus N` over NCCL up
to the transport underneath torch.distributed.
"""
import os

import numpy as np
import pytest
from numpy.testing import assert_allclose

pytestmark = pytest.mark.gpu

native = pytest.importorskip("gaussian_process_optimization_b200.native")
from gaussian_process_optimization_b200 import GPy, _lib, sharded  # noqa: E402
from gaussian_process_optimization_b200.models import StalePosteriorError  # noqa: E402


def _data(n=200, d=3, seed=3):
    rs = np.random.RandomState(seed)
    X = rs.uniform(0, 1, (n, d))
    Y = np.sin(X.sum(1))[:, None] + 0.05 * rs.randn(n, 1)
    return X, (Y - Y.mean()) / Y.std()


def test_posterior_is_a_stamped_view_of_the_resident_model():
    X, Y = _data()
    m = GPy.models.GPRegression(X, Y, kernel=GPy.kern.RBF(3, ARD=True), noise_var=0.1)
    old = m.posterior
    L_old = old.woodbury_chol.copy()               # fetched while current: a host snapshot
    mu_old, _ = m.predict(X[:5])
    m.kern.variance[:] = 2.0                        # parameters_changed -> a new fit, a new posterior
    assert m.posterior is not old
    with pytest.raises(StalePosteriorError):
        old._raw_predict(m.kern, X[:5], X)
    with pytest.raises(StalePosteriorError):
        old.woodbury_inv                            # never fetched while current: would read the NEW model's Ky^-1
    assert np.array_equal(old.woodbury_chol, L_old)  # the snapshot taken in time is still served
    mu_new, _ = m.predict(X[:5])
    assert not np.allclose(mu_new, mu_old)


def test_posterior_after_a_failed_fit_is_stale_not_silently_wrong():
    X, Y = _data(60, 2)
    X[1] = X[0]                                     # duplicated input and (below) zero noise: not positive definite even with jitter
    m = GPy.models.GPRegression(X, Y, kernel=GPy.kern.RBF(2), noise_var=0.1)
    good = m.posterior
    with pytest.raises(np.linalg.LinAlgError):
        m.kern.lengthscale[:] = np.nan
    with pytest.raises(StalePosteriorError):
        good._raw_predict(m.kern, X[:3], X)
    m.kern.lengthscale[:] = 0.7                     # a valid write refits; the model's own posterior works again
    mu, var = m.predict(X[:3])
    assert np.all(np.isfinite(mu)) and np.all(var > 0)


def test_topk_slots_without_a_finite_score_are_reported_empty():
    X, Y = _data()
    nm = native.NativeModel("rbf", True, 3, 1, n_cap=256, cand_block=128)
    nm.set_data(X, Y)
    nm.set_theta(1.0, np.array([0.5, 0.6, 0.7]), 1e-2)
    assert nm.fit(False)[0] == 0
    fmin = nm.fmin()
    Xc = np.random.RandomState(1).uniform(0, 1, (8, 3))
    Xc[2:] = np.nan                                 # only two candidates have a finite score
    vals, idx, pts = nm.acq_topk("EI", 0.01, fmin, Xc, 5)
    assert sorted(idx[:2].tolist()) == [0, 1] and np.all(idx[2:] == -1)
    assert np.all(np.isfinite(vals[:2])) and np.all(np.isnan(vals[2:])) and np.all(np.isnan(pts[2:]))
    mv, mi, mp = sharded.merge_topk(vals, idx, pts, 5)
    assert mi.tolist() == idx[:2].tolist()
    import torch
    rows, _, _ = nm.acq_topk_dev("EI", 0.01, fmin, torch.from_numpy(Xc).cuda(), 5)
    torch.cuda.synchronize()
    r = rows.cpu().numpy()
    assert r[:2, 1].tolist() == idx[:2].astype(float).tolist() and np.all(r[2:, 1] == -1) and np.all(np.isnan(r[2:, 0]))
    assert np.array_equal(r[:2, 0], vals[:2]) and np.array_equal(r[:2, 2:], Xc[idx[:2]])
    nm.close()


def test_input_dimension_cap_is_enforced_at_creation():
    with pytest.raises(_lib.GpbError, match="input_dim"):
        native.NativeModel("rbf", True, 65, 1, n_cap=128, cand_block=128)
    nm = native.NativeModel("rbf", True, 64, 1, n_cap=128, cand_block=128)       # the advertised maximum: fit AND query work
    rs = np.random.RandomState(0)
    X = rs.uniform(0, 1, (100, 64))
    Y = rs.randn(100, 1)
    nm.set_data(X, Y)
    nm.set_theta(1.0, np.full(64, 2.0), 0.1)
    assert nm.fit(True)[0] == 0
    r = nm.acquisition("EI", 0.01, nm.fmin(), X[:3] + 0.01, with_gradients=True)
    assert np.all(np.isfinite(r["f"])) and np.all(np.isfinite(r["df"]))
    nm.close()


# ---------------------------------------------------------------------------------------------------------------------
def _rank_worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    # one GPU per rank over NCCL when the box has them (the exchange then runs device to device over NVLink); both ranks on GPU 0 over
    # gloo on the driver's single-GPU test box (NCCL refuses two ranks on one device)
    many = torch.cuda.device_count() >= world
    dist.init_process_group("nccl" if many else "gloo", rank=rank, world_size=world)
    torch.cuda.set_device(rank if many else 0)
    n, d = 700, 5
    X, Y = _data(n, d, seed=11)
    ls = 0.4 + 0.1 * np.arange(d)
    nm = native.NativeModel("mat52", True, d, 1, n_cap=n, cand_block=256)
    nm.set_data(X, Y)
    if rank == 0:                                   # only rank 0 fits
        nm.set_theta(1.3, ls, 1e-3)
        info, logL, _ = nm.fit(False)
        assert info == 0
    nm.broadcast_state(src=0)                       # the others adopt theta, alpha, L^-1
    fmin = nm.fmin()
    Xc = np.random.RandomState(5).uniform(0, 1, (3000, d))
    lo, hi = sharded.divide_candidates(Xc.shape[0], rank, world)
    rows, f, df = nm.acq_topk_dev("EI", 0.01, fmin, torch.from_numpy(Xc[lo:hi]).cuda(), 5, index_offset=lo, with_gradients=True)
    vals, idx, pts = sharded.all_gather_topk_device(rows, 5)
    if rank != 0:                                   # L was not part of the broadcast: asking for it is refused, not answered wrongly
        with pytest.raises(_lib.GpbError):
            nm.get("L")
    np.savez(out % rank, vals=vals, idx=idx, pts=pts, fmin=fmin, f=f.cpu().numpy().ravel(), df=df.cpu().numpy(), lo=lo, hi=hi,
             alpha=nm.get("alpha"))
    nm.close()
    dist.destroy_process_group()


def test_two_ranks_broadcast_the_fitted_state_and_gather_anchors_on_the_device(tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / "r%d.npz")
    mp.spawn(_rank_worker, args=(2, 29800 + (os.getpid() % 2000), out), nprocs=2, join=True)
    n, d = 700, 5
    X, Y = _data(n, d, seed=11)
    nm = native.NativeModel("mat52", True, d, 1, n_cap=n, cand_block=256)
    nm.set_data(X, Y)
    nm.set_theta(1.3, 0.4 + 0.1 * np.arange(d), 1e-3)
    assert nm.fit(False)[0] == 0
    fmin = nm.fmin()
    Xc = np.random.RandomState(5).uniform(0, 1, (3000, d))
    vals, idx, pts, f, df = nm.acq_topk_full("EI", 0.01, fmin, Xc, 5)
    z = [np.load(out % r) for r in range(2)]
    for r in range(2):
        assert np.array_equal(z[r]["idx"], idx) and np.array_equal(z[r]["vals"], vals) and np.array_equal(z[r]["pts"], pts)
        assert z[r]["fmin"] == fmin
        assert np.array_equal(z[r]["alpha"], nm.get("alpha"))            # the adopted state IS rank 0's fit, bit for bit
        lo, hi = int(z[r]["lo"]), int(z[r]["hi"])
        assert_allclose(z[r]["f"], f.ravel()[lo:hi], rtol=1e-13, atol=1e-300)   # (block boundaries differ between shard and full pass)
        assert_allclose(z[r]["df"], df[lo:hi], rtol=1e-11, atol=1e-13 * np.abs(df).max())
    nm.close()


@pytest.mark.parametrize("n,d", [(700, 5), (5000, 16)])
def test_rows_of_a_batched_call_equal_the_single_row_calls_bitwise(n, d):
    """What LockstepEvaluator relies on: up to 8 candidates share one pass over the triangle of L^-1 and each one's sums are the ones
    the M = 1 call forms (EI, LCB and the penalised LP acquisition, value and gradient)."""
    X, Y = _data(n, d, seed=n)
    nm = native.NativeModel("mat52", True, d, 1, n_cap=n, cand_block=256)
    nm.set_data(X, Y)
    nm.set_theta(1.1, 0.4 + 0.05 * np.arange(d), 1e-3)
    assert nm.fit(False)[0] == 0
    fmin = nm.fmin()
    rs = np.random.RandomState(8)
    Xb = rs.uniform(0, 1, (3, d))
    for mc in (2, 3, 5, 8):
        Xc = rs.uniform(0, 1, (mc, d))
        for acq, par in (("EI", 0.01), ("LCB", 2.0)):
            # LCB takes the softplus transform, like the reference (LP.py:31-34: log of a negative acquisition is NaN otherwise)
            nm.set_penalizers("none" if acq == "EI" else "softplus", Xb, np.array([0.3, 0.2, 0.1]), np.array([0.05, 0.04, 0.03]))
            r = nm.acquisition(acq, par, fmin, Xc, with_gradients=True, want_moments=True)
            f_lp, df_lp = nm.acquisition_lp(acq, par, fmin, Xc, with_gradients=True)
            for i in range(mc):
                r1 = nm.acquisition(acq, par, fmin, Xc[i:i + 1], with_gradients=True, want_moments=True)
                for key in ("f", "df", "m", "s", "dmdx", "dsdx"):
                    assert np.array_equal(r[key][i:i + 1], r1[key]), (mc, acq, key)
                f1, df1 = nm.acquisition_lp(acq, par, fmin, Xc[i:i + 1], with_gradients=True)
                # (an EI that underflows to 0 makes the reference's 1 / acq scale infinite, LP.py:121-126: inf / nan must match too)
                assert np.array_equal(f_lp[i:i + 1], f1, equal_nan=True) and np.array_equal(df_lp[i:i + 1], df1, equal_nan=True)
    nm.close()


def test_bo_anchor_refinement_in_lockstep_equals_the_sequential_loop():
    """AcquisitionOptimizer on the CUDA GPModel refines its 5 anchors concurrently with coalesced device calls; the suggested point
    must be the one the sequential loop finds, bit for bit, in fewer device calls."""
    from gaussian_process_optimization_b200 import GPyOpt
    rs = np.random.RandomState(2)
    X = rs.uniform(0, 1, (60, 4))
    Y = (np.sin(3 * X[:, :1]) + (X[:, 1:2] - 0.4) ** 2 + 0.1 * X[:, 2:3] * X[:, 3:4])
    space = GPyOpt.core.task.space.Design_space([{'name': 'x%d' % i, 'type': 'continuous', 'domain': (0, 1)} for i in range(4)])
    out = {}
    for mode in (True, False):
        gm = GPyOpt.models.GPModel(exact_feval=True, optimize_restarts=1, verbose=False, max_iters=50)
        np.random.seed(5)
        gm.updateModel(X, Y, None, None)
        opt = GPyOpt.optimization.AcquisitionOptimizer(space, lockstep_anchors=mode)
        acq = GPyOpt.acquisitions.AcquisitionEI(gm, space, opt, jitter=0.01)
        np.random.seed(7)
        c0 = native.launch_count()
        x, fx = acq.optimize()
        out[mode] = (x, fx, native.launch_count() - c0, getattr(opt, "lockstep_stats", None))
    assert np.array_equal(out[True][0], out[False][0]) and np.array_equal(out[True][1], out[False][1])
    assert out[True][3] is not None and out[True][3]["device_calls"] < out[True][3]["requests"]
    assert out[True][2] < out[False][2]                   # fewer kernels launched for the same answer


@pytest.mark.parametrize("engine", [0, 18])
def test_two_lane_block_pipeline_equals_the_single_lane_pass(engine):
    """Device-resident candidates in several blocks alternate between two buffer sets / streams (block b + 1's covariance rows and small
    kernels run underneath block b's GEMMs); values, gradients and anchors must be bit-identical to the host-buffer pass, which walks the
    blocks one after the other on one stream.  Also with the int8 engine on (its plane workspaces are per stream)."""
    import torch
    n, d = 1500, 6
    X, Y = _data(n, d, seed=21)
    native.set_ozaki(256 if engine else 0, engine if engine else 8)
    try:
        nm = native.NativeModel("mat52", True, d, 1, n_cap=n, cand_block=1024)
        nm.set_data(X, Y)
        nm.set_theta(1.2, 0.4 + 0.05 * np.arange(d), 1e-3)
        assert nm.fit(False)[0] == 0
        fmin = nm.fmin()
        Xc = np.random.RandomState(9).uniform(0, 1, (1024 * 4 + 300, d))          # five blocks, the last one ragged
        v0, i0, p0, f0, df0 = nm.acq_topk_full("EI", 0.01, fmin, Xc, 5)            # host buffers: one lane
        Xd = torch.from_numpy(Xc).cuda()
        for rep in range(2):
            v1, i1, p1, f1, df1 = nm.acq_topk_full("EI", 0.01, fmin, Xd, 5)        # device buffers: two lanes
            assert np.array_equal(i0, i1) and np.array_equal(v0, v1) and np.array_equal(p0, p1)
            assert np.array_equal(f0, f1.cpu().numpy()) and np.array_equal(df0, df1.cpu().numpy())
            rows, f2, df2 = nm.acq_topk_dev("EI", 0.01, fmin, Xd, 5, with_gradients=True)
            torch.cuda.synchronize()
            r = rows.cpu().numpy()
            assert np.array_equal(r[:, 0], v0) and np.array_equal(r[:, 1], i0.astype(float)) and np.array_equal(r[:, 2:], p0)
            assert np.array_equal(f2.cpu().numpy(), f0) and np.array_equal(df2.cpu().numpy(), df0)
        nm.close()
    finally:
        native.set_ozaki(0)


def test_gower_model_two_lane_pass_with_a_short_tail_block_equals_the_host_buffer_pass():
    """A Gower model takes the multi-kernel route for a tail block of <= 8 candidates.  In the two-lane scoring pass that block runs
    on the second lane's stream; its triangular products used to be issued on the factor's own stream, unordered against the lane's
    covariance rows (found by the random sweep with GPB_SKINNY_FUSED=0).  Device-buffer pass (two lanes) == host-buffer pass, bitwise."""
    import torch
    rs = np.random.RandomState(21)
    n, d = 300, 4
    X = rs.uniform(0, 1, (n, d))
    X[:, 3] = rs.randint(0, 3, n)
    Y = np.sin(3 * X.sum(1))[:, None] + 0.05 * rs.randn(n, 1)
    nm = native.NativeModel("mat52", True, d, 1, n_cap=n, cand_block=128)
    nm.set_data(X, Y)
    nm.set_gower(([0, 1, 2], [3], [1.0, 1.5, 0.7]))
    nm.set_theta(1.2, np.ones(d), 1e-2)
    assert nm.fit(False)[0] == 0
    fmin = nm.fmin()
    for M in (128 * 2 + 3, 128 * 3 + 8, 128 + 1):
        Xc = rs.uniform(0, 1, (M, d))
        Xc[:, 3] = rs.randint(0, 3, M)
        for rep in range(3):
            a = nm.acq_topk_full("EI", 0.01, fmin, Xc, 5)
            b = nm.acq_topk_full("EI", 0.01, fmin, torch.from_numpy(Xc).cuda(), 5)
            for i in range(3):
                assert np.array_equal(a[i], b[i]), (M, rep, i)
            assert np.array_equal(a[3], b[3].cpu().numpy()) and np.array_equal(a[4], b[4].cpu().numpy()), (M, rep)
    nm.close()
