"""CPU tests of the host-side mirror: parameter transforms / optimiser-array plumbing, design space, RNG consumption order of
the random design (trajectory parity depends on it), normalisation, and the sharded top-k collective (gloo, world size 2)."""
import os
import sys

import numpy as np
import pytest
from numpy.testing import assert_allclose

from gaussian_process_optimization_b200 import gpyopt as G
from gaussian_process_optimization_b200 import parameterization as P
from gaussian_process_optimization_b200 import sharded
from oracle import gp_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_transforms_match_oracle_restatement():
    x = np.array([-800., -40., -1., 0., 0.5, 3., 35.9, 36.1, 50.])
    f = P.Logexp().f(x)
    assert_allclose(f, O.Logexp.f(x), rtol=0, atol=0)
    assert np.all(f >= 0) and np.all(np.isfinite(f))
    pos = np.array([1e-9, 1e-3, 0.5, 1., 10., 37., 100.])
    assert_allclose(P.Logexp().f(P.Logexp().finv(pos)), pos, rtol=1e-12)
    assert_allclose(P.Logexp().gradfactor(pos, np.ones_like(pos)), O.Logexp.gradfactor(pos, np.ones_like(pos)))
    lg, lo = P.Logistic(1e-9, 1e6), O.Logistic(1e-9, 1e6)
    v = np.array([1e-8, 1e-3, 1., 1e5])
    assert_allclose(lg.finv(v), lo.finv(v))
    assert_allclose(lg.f(lg.finv(v)), v, rtol=1e-9)
    assert_allclose(lg.gradfactor(v, 2 * np.ones(4)), lo.gradfactor(v, 2 * np.ones(4)))
    # gradfactor is d theta / d x
    x0 = np.array([0.3])
    num = (P.Logexp().f(x0 + 1e-6) - P.Logexp().f(x0 - 1e-6)) / 2e-6
    assert_allclose(P.Logexp().gradfactor(P.Logexp().f(x0), 1.0), num, rtol=1e-8)


class _Toy(P.Model):
    """log_likelihood = -sum((theta - t)^2) with a positive, a bounded and a fixed parameter."""

    def __init__(self):
        super(_Toy, self).__init__("toy")
        self.calls = 0
        self.t = np.array([0.7, 1.5, 1.2, 9.9])
        self.a = P.Param("a", [1.0, 2.0], P.Logexp())
        self.b = P.Param("b", 0.5)
        self.c = P.Param("c", 3.0, P.Logexp())
        self.link_parameters(self.a, self.b, self.c)
        self.b.constrain_bounded(0.0, 2.0)
        self.c.constrain_fixed(4.0)

    def parameters_changed(self):
        self.calls += 1
        th = self.param_array
        self._ll = -np.sum((th - self.t) ** 2)
        g = -2 * (th - self.t)
        self.a.gradient, self.b.gradient, self.c.gradient = g[:2], g[2], g[3]

    def log_likelihood(self):
        return self._ll


def test_model_parameter_plumbing_and_optimize():
    m = _Toy()
    m.parameters_changed()
    assert m.size == 4 and m._size_transformed() == 3
    assert list(m.parameter_names_flat()) == ["toy.a[[0]]", "toy.a[[1]]", "toy.b"]
    assert list(m.parameter_names_flat(include_fixed=True))[-1] == "toy.c"
    assert_allclose(m[:], [1., 2., .5, 4.])
    x = m.optimizer_array.copy()
    n0 = m.calls
    m.optimizer_array = x + 0.1
    assert m.calls == n0 + 1                       # one parameters_changed per write
    assert m.param_array[3] == 4.0                 # fixed parameter untouched
    m.a = [1.1, 2.2]                               # attribute write goes through the observer
    assert m.calls == n0 + 2
    assert_allclose(m.a.values, [1.1, 2.2])
    m.update_model(False)
    m.a[0] = 3.0
    assert m.calls == n0 + 2
    m.update_model(True)
    assert m.calls == n0 + 3
    assert m.checkgrad()
    run = m.optimize(max_iters=200)
    assert_allclose(m.param_array[:3], m.t[:3], atol=1e-4)
    assert run.status == "Converged"
    np.random.seed(1)
    runs = m.optimize_restarts(num_restarts=3, verbose=False)
    assert len(runs) == 4
    assert_allclose(m.param_array[:3], m.t[:3], atol=1e-4)


def test_randomize_draws_one_normal_vector():
    m = _Toy()
    np.random.seed(7)
    expect = np.random.normal(size=3)
    after = np.random.uniform()
    np.random.seed(7)
    m.randomize()
    assert_allclose(m.optimizer_array, expect, rtol=1e-12)
    assert np.random.uniform() == after            # exactly one draw of n_free normals was consumed


def test_design_space_and_random_design_rng_order():
    space = G.Design_space([{'name': 'x', 'type': 'continuous', 'domain': (-5, 10)},
                            {'name': 'k', 'type': 'discrete', 'domain': (0, 1, 2, 5)},
                            {'name': 'y', 'type': 'continuous', 'domain': (1, 15), 'dimensionality': 2}])
    assert space.dimensionality == 4
    assert space.get_bounds() == [(-5, 10), (0, 5), (1, 15), (1, 15)]
    assert space.get_continuous_dims() == [0, 2, 3]
    np.random.seed(3)
    X = G.initial_design('random', space, 6)
    np.random.seed(3)                              # discrete first, then one uniform(size=n) per continuous dimension
    k = np.random.choice((0, 1, 2, 5), 6)
    c0 = np.random.uniform(-5, 10, 6)
    c1 = np.random.uniform(1, 15, 6)
    c2 = np.random.uniform(1, 15, 6)
    assert_allclose(X, np.stack([c0, k, c1, c2], 1))
    r = space.round_optimum(np.array([[11.0, 3.4, 0.0, 7.0]]))
    assert_allclose(r, [[10.0, 2.0, 1.0, 7.0]])
    cs = G.Design_space(G.bounds_to_space([(0, 1), (0, 1)]), [{'name': 'c', 'constraint': 'x[:,0] + x[:,1] - 1'}])
    assert_allclose(cs.indicator_constraints(np.array([[0.2, 0.2], [0.9, 0.9]])), [[1], [0]])


def test_normalize_and_best_value():
    Y = np.array([[3.], [1.], [2.]])
    assert_allclose(G.normalize(Y), (Y - 2.) / Y.std())
    assert_allclose(G.normalize(np.ones((3, 1))), np.zeros((3, 1)))
    assert_allclose(G.best_value(Y), [3., 1., 1.])


def test_anchor_generator_and_lbfgs_on_a_quadratic():
    space = G.Design_space(G.bounds_to_space([(-2, 2), (-2, 2)]))

    def f(x):
        x = np.atleast_2d(x)
        return np.sum((x - 0.5) ** 2, 1)[:, None]

    def f_df(x):
        x = np.atleast_2d(x)
        return f(x), 2 * (x - 0.5)

    np.random.seed(0)
    opt = G.AcquisitionOptimizer(space)
    x, fx = opt.optimize(f=f, f_df=f_df)
    assert_allclose(x, [[0.5, 0.5]], atol=1e-5)
    assert fx.shape == (1, 1)


def test_divide_and_merge_topk():
    ranges = [sharded.divide_candidates(10, r, 4) for r in range(4)]
    assert ranges == [(0, 3), (3, 6), (6, 8), (8, 10)]
    v, i, p = sharded.merge_topk([0.5, 0.1, 0.1, 0.7], [4, 9, 2, 1], np.arange(8).reshape(4, 2), 3)
    assert list(i) == [2, 9, 4]
    assert_allclose(v, [0.1, 0.1, 0.5])


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rs = np.random.RandomState(11)
    X = rs.uniform(0, 1, (37, 3))
    X[5] = X[30]                                   # a tie across shards: the lowest global index must win
    score = lambda Z: np.sin(7 * Z[:, 0]) + Z[:, 1] * Z[:, 2]  # noqa: E731
    a, b = sharded.divide_candidates(X.shape[0], rank, world)
    sc = sharded.ShardedAnchorScorer(score_fn=score)
    vals, idx, pts = sc.topk(X[a:b], 5, index_offset=a)
    # the same exchange from a (k, d + 2) row buffer [f, global index, coordinates] as acq_topk_dev leaves it (the gloo branch of
    # all_gather_topk_device; one slot of rank 1 left empty: index -1 rows must not survive the merge)
    import torch
    s_loc = score(X[a:b])
    o = np.lexsort((np.arange(a, b), s_loc))[:5]
    rows = np.column_stack([s_loc[o], (a + o).astype(np.float64), X[a:b][o]])
    if rank == 1:
        rows[-1] = np.nan
        rows[-1, 1] = -1.0
    dv, di, dp = sharded.all_gather_topk_device(torch.from_numpy(rows), 5)
    np.savez(out % rank, vals=vals, idx=idx, pts=pts, dvals=dv, didx=di, dpts=dp)
    dist.destroy_process_group()


def test_sharded_topk_gloo_world_size_2(tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / "r%d.npz")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    rs = np.random.RandomState(11)
    X = rs.uniform(0, 1, (37, 3))
    X[5] = X[30]
    s = np.sin(7 * X[:, 0]) + X[:, 1] * X[:, 2]
    order = np.argsort(s, kind="stable")[:5]
    for r in range(2):
        z = np.load(out % r)
        assert np.array_equal(z["idx"], order)
        assert_allclose(z["vals"], s[order])
        assert_allclose(z["pts"], X[order])
        assert np.array_equal(z["didx"], order)
        assert_allclose(z["dvals"], s[order])
        assert_allclose(z["dpts"], X[order])


# ---------------------------------------------------------------------------------------------------------------------
# LockstepEvaluator: concurrent L-BFGS-B runs, their f_df requests coalesced into batched calls
# ---------------------------------------------------------------------------------------------------------------------
def test_lockstep_evaluator_gives_every_run_its_sequential_trajectory():
    """Five bounded L-BFGS-B runs on a row-wise function: in lockstep (threads + coalesced calls) every run must see exactly the
    values the sequential loop gives it -- same iterates, same optimum -- while the number of batched calls is the LONGEST run's
    request count, not the sum."""
    from gaussian_process_optimization_b200 import gpyopt as G

    class Acq(object):
        batched_rows_bitwise = True

        def __init__(self):
            self.calls, self.rows = 0, 0

        def f(self, X):
            X = np.atleast_2d(X)
            return (np.sum((X - 0.3) ** 2 * np.array([1.0, 10.0, 100.0]), axis=1) + np.sin(3 * X[:, 0]))[:, None]

        def f_df(self, X):
            X = np.atleast_2d(X)
            self.calls += 1
            self.rows += X.shape[0]
            df = 2 * (X - 0.3) * np.array([1.0, 10.0, 100.0])
            df[:, 0] += 3 * np.cos(3 * X[:, 0])
            return self.f(X), df

    space = G.Design_space([{'name': 'x%d' % i, 'type': 'continuous', 'domain': (0, 1)} for i in range(3)])
    anchors = np.random.RandomState(3).uniform(0, 1, (5, 3))
    a_seq, a_lock = Acq(), Acq()
    opt = G.AcquisitionOptimizer(space)
    opt.optimizer = G.OptLbfgs(space.get_bounds())
    seq = [G.apply_optimizer(opt.optimizer, a, f=a_seq.f, f_df=a_seq.f_df, space=space) for a in anchors]
    lock = opt._optimize_anchors_lockstep(anchors, a_lock.f, a_lock.f_df, None)
    for (xs, fs), (xl, fl) in zip(seq, lock):
        assert np.array_equal(xs, xl) and np.array_equal(fs, fl)
    assert a_lock.rows == a_seq.rows                      # the same requests were answered ...
    assert a_lock.calls < a_seq.calls                     # ... in fewer calls
    assert opt.lockstep_stats["requests"] == a_seq.rows
    opt.kwargs['lockstep_anchors'] = True
    assert opt._lockstep_ok(a_lock.f_df, anchors) and not opt._lockstep_ok(a_lock.f_df, anchors[:1])
    opt.kwargs['lockstep_anchors'] = 'auto'            # automatic mode: only for models of at least LOCKSTEP_MIN_N points
    assert not opt._lockstep_ok(a_lock.f_df, anchors)
    a_lock.model = type("M", (), {"model": type("G", (), {"X": np.zeros((opt.LOCKSTEP_MIN_N, 3))})()})()
    assert opt._lockstep_ok(a_lock.f_df, anchors)
    opt.kwargs['lockstep_anchors'] = False
    assert not opt._lockstep_ok(a_lock.f_df, anchors)


def test_lockstep_evaluator_propagates_errors_and_does_not_deadlock():
    from gaussian_process_optimization_b200 import gpyopt as G

    def bad(X):
        raise RuntimeError("device call failed")

    space = G.Design_space([{'name': 'x', 'type': 'continuous', 'domain': (0, 1)}])
    opt = G.AcquisitionOptimizer(space)
    opt.optimizer = G.OptLbfgs(space.get_bounds())
    with pytest.raises(RuntimeError, match="device call failed"):
        opt._optimize_anchors_lockstep(np.array([[0.2], [0.7], [0.9]]), lambda X: np.zeros((np.atleast_2d(X).shape[0], 1)), bad, None)
