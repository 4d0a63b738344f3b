"""The GPy / GPyOpt-shaped host API (`from gaussian_process_optimization_b200 import GPy, GPyOpt`).

Two backends run the same test bodies:
  * "cuda"   (marked gpu): GPRegression / GPModel on libgpb200.so -- the product path;
  * "oracle" (CPU):        the same host classes on the CPU oracle (tests/oracle_backend.py) -- this validates the host logic
                           (parameter plumbing, transforms, L-BFGS-B, BO loop) without a GPU and gives the cuda run its
                           comparison target for optimiser runs and BO trajectories.
The bodies re-express the reference's own tests for this path: GradientTests (model_tests.py:647-723),
test_raw_predict (:63-82), test_raw_predict_numerical_stability (:25-61), test_setxy_gp (gp_tests.py:50-60),
jitchol success / failure (linalg_test.py:7-37), and config 1 of BASELINE.json (BO on Branin).
"""
import numpy as np
import pytest
from numpy.testing import assert_allclose

from gaussian_process_optimization_b200 import GPy, GPyOpt, _lib
from oracle import gp_oracle as O
import oracle_backend as OB


def _has_gpu():
    try:
        return _lib.load().gpb_device_count() > 0
    except Exception:
        return False


BACKENDS = [pytest.param("cuda", marks=pytest.mark.gpu), "oracle"]


def make_gpr(backend, X, Y, kernel, noise_var=1.):
    if backend == "cuda":
        return GPy.models.GPRegression(X, Y, kernel=kernel, noise_var=noise_var)
    return OB.oracle_gp_regression(X, Y, kernel, noise_var)


def make_gpmodel(backend, **kw):
    return GPyOpt.models.GPModel(**kw) if backend == "cuda" else OB.OracleGPModel(**kw)


def branin(X):
    """GPyOpt/GPyOpt/objective_examples/experiments2d.py:203-216 (sd = 0)."""
    X = np.atleast_2d(X)
    x1, x2 = X[:, 0], X[:, 1]
    b, c, r, s, t = 5.1 / (4 * np.pi ** 2), 5 / np.pi, 6, 10, 1 / (8 * np.pi)
    return ((x2 - b * x1 ** 2 + c * x1 - r) ** 2 + s * (1 - t) * np.cos(x1) + s).reshape(-1, 1)


BRANIN_DOMAIN = [{'name': 'x1', 'type': 'continuous', 'domain': (-5, 10)}, {'name': 'x2', 'type': 'continuous', 'domain': (1, 15)}]


# ---------------------------------------------------------------------------------------------------------------------
# GradientTests.check_model (model_tests.py:647-723): randomize + checkgrad for rbf / matern52, ARD or not, 1-D and 2-D
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("kname,dim,ard", [("RBF", 1, False), ("RBF", 2, False), ("RBF", 2, True), ("Matern52", 1, False),
                                           ("Matern52", 2, False), ("Matern52", 2, True)])
def test_gpregression_checkgrad(backend, kname, dim, ard):
    np.random.seed(11 + dim)
    if dim == 1:
        X = np.random.uniform(-3., 3., (20, 1))
        Y = np.sin(X) + np.random.randn(20, 1) * 0.05
    else:
        X = np.random.uniform(-3., 3., (40, 2))
        Y = np.sin(X[:, 0:1]) * np.sin(X[:, 1:2]) + np.random.randn(40, 1) * 0.05
    k = getattr(GPy.kern, kname)(dim, ARD=ard)
    m = make_gpr(backend, X, Y, k)
    m.randomize()
    assert m.checkgrad()
    # parameter order in m[:] (stationary.py:83, core/gp.py:108-109)
    names = m.parameter_names_flat()
    assert names[0].endswith("variance") and names[-1].endswith("Gaussian_noise.variance")
    assert m[:].size == 2 + (dim if ard else 1)


@pytest.mark.parametrize("backend", BACKENDS)
def test_raw_predict_closed_form(backend):
    """model_tests.py:63-82: predict_noiseless against the explicit pinv(K + s2 I) formula, 7 decimals."""
    np.random.seed(12345)
    N, N_new, D = 20, 50, 1
    X = np.random.uniform(-3., 3., (N, 1))
    Y = np.sin(X) + np.random.randn(N, D) * 0.05
    X_new = np.random.uniform(-3., 3., (N_new, 1))
    k = GPy.kern.RBF(1)
    m = make_gpr(backend, X, Y, k)
    m.randomize()
    m.likelihood.variance = .5
    v, ls = float(k.variance.values[0]), k.lengthscale.values
    Kf = lambda A, B=None: O.K("rbf", A, B, v, ls, False)  # noqa: E731  (closed form on the host, independent of the backend)
    Kinv = np.linalg.pinv(Kf(X) + np.eye(N) * 0.5)
    K_hat = Kf(X_new) - Kf(X_new, X).dot(Kinv).dot(Kf(X, X_new))
    mu_hat = Kf(X_new, X).dot(Kinv).dot(m.Y_normalized)
    mu, covar = m.predict_noiseless(X_new, full_cov=True)
    assert mu.shape == (N_new, D) and covar.shape == (N_new, N_new)
    np.testing.assert_almost_equal(K_hat, covar)
    np.testing.assert_almost_equal(mu_hat, mu)
    mu, var = m.predict_noiseless(X_new)
    assert mu.shape == (N_new, D) and var.shape == (N_new, 1)
    np.testing.assert_almost_equal(np.diag(K_hat)[:, None], var)
    np.testing.assert_almost_equal(mu_hat, mu)


@pytest.mark.parametrize("backend", BACKENDS)
def test_setxy_roundtrip(backend):
    """gp_tests.py:50-60."""
    np.random.seed(12345)
    X = np.random.uniform(-3., 3., (20, 1))
    Y = np.sin(X) + np.random.randn(20, 1) * 0.05
    m = make_gpr(backend, X, Y, GPy.kern.RBF(1))
    mu, var = m.predict(m.X)
    Xc = m.X.copy()
    m.set_XY(m.X[:10], m.Y[:10])
    assert m.checkgrad()
    m.set_XY(Xc, Y)
    mu2, var2 = m.predict(m.X)
    assert_allclose(mu, mu2)
    assert_allclose(var, var2)


@pytest.mark.parametrize("backend", BACKENDS)
def test_branin_grid_variance_nonnegative(backend):
    """model_tests.py:25-61: 5x5 Branin grid, RBF-ARD, noise fixed 1e-5, randomize + optimize, var >= 0 at 1e5 points."""
    np.random.seed(3)
    xg1, xg2 = np.linspace(-5, 10, 5), np.linspace(0, 15, 5)
    X = np.zeros((25, 2))
    for i, x1 in enumerate(xg1):
        for j, x2 in enumerate(xg2):
            X[i + 5 * j, :] = [x1, x2]
    Y = branin(X)
    m = make_gpr(backend, X, Y, GPy.kern.RBF(input_dim=2, ARD=True))
    m.likelihood.variance.fix(1e-5)
    m.randomize()
    m.optimize()
    Xp = np.random.uniform(size=(int(1e5) if backend == "cuda" else 5000, 2))
    Xp[:, 0] = Xp[:, 0] * 15 - 5
    Xp[:, 1] = Xp[:, 1] * 15
    _, var = m.predict(Xp)
    assert np.all(var >= 0.)
    assert m.parameter_names_flat().size == 3      # the fixed noise is not an optimiser variable


# ---------------------------------------------------------------------------------------------------------------------
# GPyOpt GPModel / acquisitions through the public classes, against the golden vectors of the reference's own sources
# ---------------------------------------------------------------------------------------------------------------------
def _cond_tol(g):
    K = g["K"]
    w = np.linalg.eigvalsh(K + (g["noise"] + 1e-8) * np.eye(K.shape[0]))
    return max(1.0, w[-1] / w[0] * 2.2e-16 / 1e-12)


@pytest.mark.parametrize("backend", BACKENDS)
def test_public_classes_golden(backend, golden):
    g = golden
    D = g["X"].shape[1]
    ct = _cond_tol(g)
    K = GPy.kern.RBF if g["kind"] == "rbf" else GPy.kern.Matern52
    k = K(D, variance=g["variance"], lengthscale=g["lengthscale"], ARD=g["ard"])
    m = make_gpr(backend, g["X"], g["Y"], k, noise_var=g["noise"])
    assert_allclose(m.log_likelihood(), g["logL"], rtol=1e-9 * ct)
    assert_allclose(np.ravel(k.variance.gradient), g["grad_var"].ravel(), rtol=1e-7 * ct)
    assert_allclose(np.ravel(k.lengthscale.gradient), g["grad_len"].ravel(), rtol=1e-7 * ct)
    assert_allclose(np.ravel(m.likelihood.variance.gradient), g["grad_noise"].ravel(), rtol=1e-7 * ct)
    assert_allclose(m.posterior.woodbury_chol, g["L"], rtol=1e-9, atol=1e-12)
    assert_allclose(m.posterior.woodbury_vector, g["alpha"], rtol=1e-9 * ct, atol=1e-9 * ct * np.abs(g["alpha"]).max())
    mu, var = m.predict(g["Xs"])
    assert_allclose(mu, g["pred_mu"], rtol=1e-9 * ct, atol=1e-11 * ct)
    assert_allclose(var, g["pred_var"], rtol=1e-9 * ct, atol=1e-13 * ct)
    mu, cov = m.predict(g["Xs"], full_cov=True)
    assert_allclose(cov, g["pred_cov"], rtol=1e-9 * ct, atol=1e-11 * ct)
    dm, dv = m.predictive_gradients(g["Xs"])
    assert dm.shape == g["dmu_dX"].shape and dv.shape == g["dv_dX"].shape
    assert_allclose(dm, g["dmu_dX"], rtol=1e-7 * ct, atol=1e-9 * ct * np.abs(g["dmu_dX"]).max())
    assert_allclose(dv, g["dv_dX"], rtol=1e-7 * ct, atol=1e-9 * ct * np.abs(g["dv_dX"]).max())
    # GPyOpt adaptor + acquisitions (the object layout GPyOpt builds in BayesianOptimization.__init__)
    gm = make_gpmodel(backend, exact_feval=False, verbose=False)
    gm.model = m
    mm, ss = gm.predict(g["Xs"])
    assert_allclose(mm, g["gpm_m"], rtol=1e-9 * ct, atol=1e-11 * ct)
    assert_allclose(ss, g["gpm_s"], rtol=1e-9 * ct)
    mm, ss, dmdx, dsdx = gm.predict_withGradients(g["Xs"])
    assert_allclose(dmdx, g["gpm_dmdx"], rtol=1e-7 * ct, atol=1e-9 * ct * np.abs(g["gpm_dmdx"]).max())
    assert_allclose(dsdx, g["gpm_dsdx"], rtol=1e-7 * ct, atol=1e-9 * ct * np.abs(g["gpm_dsdx"]).max())
    assert_allclose(gm.get_fmin(), g["fmin"], rtol=1e-9 * ct)
    space = GPyOpt.Design_space([{'name': 'x', 'type': 'continuous', 'domain': (0, 1), 'dimensionality': D}])
    ei = GPyOpt.acquisitions.AcquisitionEI(gm, space, optimizer=None, jitter=0.01)
    lcb = GPyOpt.acquisitions.AcquisitionLCB(gm, space, optimizer=None, exploration_weight=2)
    assert_allclose(ei.acquisition_function(g["Xs"]), g["ei"], rtol=1e-8 * ct, atol=1e-14)
    f, df = ei.acquisition_function_withGradients(g["Xs"])
    assert_allclose(f, g["ei_g_f"], rtol=1e-8 * ct, atol=1e-14)
    assert_allclose(df, g["ei_g_df"], rtol=1e-7 * ct, atol=1e-9 * ct * np.abs(g["ei_g_df"]).max())
    assert_allclose(lcb.acquisition_function(g["Xs"]), g["lcb"], rtol=1e-9 * ct, atol=1e-11 * ct)
    f, df = lcb.acquisition_function_withGradients(g["Xs"])
    assert_allclose(df, g["lcb_g_df"], rtol=1e-7 * ct, atol=1e-9 * ct * np.abs(g["lcb_g_df"]).max())
    # local penalisation around a batch of three points (acquisitions/LP.py), reference vectors from the reference's own LP.py
    for tag, base in (("ei", ei), ("lcb", lcb)):
        lp = GPyOpt.acquisitions.AcquisitionLP(gm, space, None, base)
        assert lp.transform == ("none" if tag == "ei" else "softplus")
        assert (lp._native_kind() is not None) == (backend == "cuda")
        lp.update_batches(g["lp_Xb"], float(g["lp_L"]), float(g["lp_Min"]))
        assert_allclose(lp.r_x0, g["lp_%s_r" % tag], rtol=1e-9 * ct, atol=1e-11 * ct)
        assert_allclose(lp.s_x0, g["lp_%s_s" % tag], rtol=1e-9 * ct)
        Xq = g["Xs"][3:]
        f_ref, df_ref = g["lp_%s_f" % tag], g["lp_%s_df" % tag]
        assert_allclose(lp.acquisition_function(Xq), f_ref, rtol=1e-7 * ct, atol=1e-7 * ct)
        got = np.vstack([lp.acquisition_function_withGradients(Xq[i:i + 1])[1] for i in range(Xq.shape[0])])
        assert_allclose(got, df_ref, rtol=1e-6 * ct, atol=1e-8 * ct * np.abs(df_ref).max())
        if backend == "cuda":     # the device path also takes many points at once
            fb, dfb = lp.acquisition_function_withGradients(Xq)
            assert_allclose(fb, f_ref, rtol=1e-7 * ct, atol=1e-7 * ct)
            assert_allclose(dfb, df_ref, rtol=1e-6 * ct, atol=1e-8 * ct * np.abs(df_ref).max())
        lp.update_batches(None, None, None)
        assert_allclose(lp.acquisition_function(Xq), g["lp_%s_f_nobatch" % tag], rtol=1e-7 * ct, atol=1e-7 * ct)
    # one point, as apply_optimizer's wrappers pass it (optimizer.py:200-232)
    f1 = ei.acquisition_function(g["Xs"][3:4])
    assert f1.shape == (1, 1)
    assert_allclose(f1[0, 0], g["ei"][3, 0], rtol=1e-8 * ct, atol=1e-14)
    m1, s1 = gm.predict(g["Xs"][3])            # GPModel itself accepts 1-D input (gpmodel.py:96-97)
    assert_allclose(m1[0, 0], g["gpm_m"][3, 0], rtol=1e-9 * ct, atol=1e-11 * ct)


# ---------------------------------------------------------------------------------------------------------------------
# jitchol ladder (linalg_test.py:7-37) -- CUDA only (the oracle's jitchol is checked in test_oracle_golden.py)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_jitchol_ladder_success_and_failure():
    np.random.seed(5)
    A = np.random.randn(20, 100)
    A = A.dot(A.T)
    vals, vectors = np.linalg.eigh(A)
    vals[vals.argmin()] = 0
    default_jitter = 1e-6 * np.mean(vals)
    vals[vals.argmin()] = -default_jitter * (10 ** 3.5)
    A_corrupt = (vectors * vals).dot(vectors.T)
    L = GPy.util.linalg.jitchol(A_corrupt, maxtries=5)
    A_new = L.dot(L.T)
    diff = A_new - A_corrupt
    assert_allclose(diff, np.eye(20) * np.diag(diff).mean(), atol=1e-11)
    # the jitter that made it through is the 5th rung: mean(diag) * 1e-6 * 10^4 (linalg.py:66-72)
    assert_allclose(np.diag(diff).mean(), np.diag(A_corrupt).mean() * 1e-6 * 1e4, rtol=1e-6)
    with pytest.raises(np.linalg.LinAlgError):
        GPy.util.linalg.jitchol(A_corrupt, maxtries=4)
    with pytest.raises(np.linalg.LinAlgError, match="non-positive diagonal"):
        GPy.util.linalg.jitchol(-np.eye(4))
    # same ladder in the oracle
    assert_allclose(O.jitchol(A_corrupt, maxtries=5), L, rtol=1e-8, atol=1e-10)


@pytest.mark.gpu
def test_linalg_surface_against_lapack():
    rs = np.random.RandomState(9)
    B = rs.randn(150, 150)
    A = B @ B.T + 150 * np.eye(150)
    Ai, L, Li, logdet = GPy.util.linalg.pdinv(A)
    Ai_r, L_r, Li_r, ld_r = O.pdinv(A)
    assert_allclose(L, L_r, rtol=1e-11, atol=1e-12)
    assert_allclose(Li, Li_r, rtol=1e-10, atol=1e-13)
    assert_allclose(Ai, Ai_r, rtol=1e-10, atol=1e-14)
    assert_allclose(logdet, ld_r, rtol=1e-13)
    Y = rs.randn(150, 2)
    assert_allclose(GPy.util.linalg.dpotrs(L, Y, lower=1)[0], O.dpotrs(L_r, Y, lower=1)[0], rtol=1e-10, atol=1e-14)
    assert_allclose(GPy.util.linalg.dpotri(L, lower=1)[0], Ai_r, rtol=1e-10, atol=1e-14)


# ---------------------------------------------------------------------------------------------------------------------
# hyper-parameter optimisation and the BO loop: CUDA run vs the identical host logic on the oracle
# ---------------------------------------------------------------------------------------------------------------------
def _opt_problem():
    rs = np.random.RandomState(21)
    X = rs.uniform(0, 1, (120, 3))
    Y = np.sin(3 * X[:, :1]) + np.cos(2 * X[:, 1:2]) * X[:, 2:3] + 0.05 * rs.randn(120, 1)
    return X, (Y - Y.mean()) / Y.std()


@pytest.mark.parametrize("backend", ["oracle"])
def test_optimize_restarts_improves_objective(backend):
    X, Y = _opt_problem()
    m = make_gpr(backend, X, Y, GPy.kern.Matern52(3, ARD=True), noise_var=0.1)
    f0 = m.objective_function()
    np.random.seed(0)
    m.optimize_restarts(num_restarts=2, optimizer='lbfgs', max_iters=200, verbose=False)
    assert m.objective_function() < f0 - 10
    assert len(m.optimization_runs) == 2
    assert_allclose(m.objective_function(), min(r.f_opt for r in m.optimization_runs), rtol=1e-12)


@pytest.mark.gpu
def test_optimize_matches_oracle_run():
    """GP.optimize (core/gp.py:643-664) from the same start: every L-BFGS-B step sees f, g equal to ~1e-11, so the two runs
    must end at the same optimum (compared at the optimiser's own tolerance, factr = 1e7 -> ~2e-9 relative in f)."""
    X, Y = _opt_problem()
    res = {}
    for backend in ("cuda", "oracle"):
        m = make_gpr(backend, X, Y, GPy.kern.Matern52(3, ARD=True), noise_var=0.1)
        m.optimize(optimizer='lbfgs', max_iters=1000)
        res[backend] = (m.objective_function(), m[:].copy(), m.optimization_runs[-1].funct_eval)
    assert_allclose(res["cuda"][0], res["oracle"][0], rtol=1e-8)
    assert_allclose(res["cuda"][1], res["oracle"][1], rtol=2e-4)
    assert res["cuda"][2] == res["oracle"][2], "different number of objective evaluations: %r" % (res,)


def _run_bo(backend, iters, seed=0, acquisition_type='EI', **kw):
    np.random.seed(seed)
    model = make_gpmodel(backend, kernel=GPy.kern.RBF(2), exact_feval=True, verbose=False, optimize_restarts=kw.pop("restarts", 5))
    bo = GPyOpt.methods.BayesianOptimization(branin, domain=BRANIN_DOMAIN, model=model, acquisition_type=acquisition_type,
                                             exact_feval=True, initial_design_numdata=5, initial_design_type='random', **kw)
    bo.run_optimization(max_iter=iters)
    return bo


@pytest.mark.parametrize("backend", BACKENDS)
def test_bo_local_penalization_batch(backend):
    """Batch BO with evaluator_type='local_penalization' (core/evaluators/batch_local_penalization.py): every iteration
    proposes batch_size points, the first from the plain log-acquisition, the others from the penalised one."""
    np.random.seed(4)
    model = make_gpmodel(backend, kernel=GPy.kern.Matern52(2), exact_feval=True, verbose=False, optimize_restarts=1)
    bo = GPyOpt.methods.BayesianOptimization(branin, domain=BRANIN_DOMAIN, model=model, acquisition_type='EI', exact_feval=True,
                                             initial_design_numdata=6, evaluator_type='local_penalization', batch_size=3)
    assert isinstance(bo.evaluator, GPyOpt.core.evaluators.LocalPenalization)
    bo.run_optimization(max_iter=3)
    assert bo.X.shape == (6 + 3 * 3, 2) and bo.Y.shape == (15, 1)
    for it in range(3):                       # the points of one batch are kept apart by the penalisers
        B = bo.X[6 + 3 * it: 9 + 3 * it]
        dist = np.sqrt(((B[:, None, :] - B[None, :, :]) ** 2).sum(-1))[np.triu_indices(3, 1)]
        assert np.all(dist > 1e-3)
    assert bo.fx_opt <= bo.Y[:6].min()


@pytest.mark.gpu
def test_bo_local_penalization_matches_oracle():
    np.random.seed(4)
    res = {}
    for backend in ("cuda", "oracle"):
        np.random.seed(4)
        model = make_gpmodel(backend, kernel=GPy.kern.Matern52(2), exact_feval=True, verbose=False, optimize_restarts=1)
        bo = GPyOpt.methods.BayesianOptimization(branin, domain=BRANIN_DOMAIN, model=model, acquisition_type='EI', exact_feval=True,
                                                 initial_design_numdata=6, evaluator_type='local_penalization', batch_size=3)
        bo.run_optimization(max_iter=2)
        res[backend] = bo.X.copy()
    assert res["cuda"].shape == res["oracle"].shape == (12, 2)
    assert_allclose(res["cuda"], res["oracle"], rtol=1e-4, atol=1e-4)


def test_bo_branin_runs_on_oracle_backend():
    """Config 1 host logic end to end on CPU: 5 random initial points + 6 EI steps; the RNG order of Appendix C holds."""
    bo = _run_bo("oracle", 6, restarts=2)
    assert bo.X.shape == (11, 2) and bo.Y.shape == (11, 1)
    np.random.seed(0)
    x1 = np.random.uniform(-5, 10, 5)      # random_design.py:73-77: one uniform(size=n) draw per dimension, in order
    x2 = np.random.uniform(1, 15, 5)
    assert_allclose(bo.X[:5], np.c_[x1, x2])
    assert bo.fx_opt == bo.Y.min() and bo.fx_opt <= bo.Y[:5].min()
    lo, hi = np.array([-5., 1.]), np.array([10., 15.])
    assert np.all(bo.X >= lo) and np.all(bo.X <= hi)
    assert bo.model_parameters_iterations.shape[1] == 3


@pytest.mark.parametrize("backend", BACKENDS)
def test_modular_bo_equals_the_wrapper(backend):
    """methods/modular_bayesian_optimization.py:24: the loop assembled from caller-built handlers proposes the same points
    as BayesianOptimization assembling them itself (same seed, same initial design)."""
    iters = 4
    ref = _run_bo(backend, iters, restarts=2)
    np.random.seed(0)
    space = GPyOpt.Design_space(BRANIN_DOMAIN)
    objective = GPyOpt.core.task.SingleObjective(branin)
    X_init = GPyOpt.experiment_design.initial_design('random', space, 5)
    model = make_gpmodel(backend, kernel=GPy.kern.RBF(2), exact_feval=True, verbose=False, optimize_restarts=2)
    aopt = GPyOpt.optimization.AcquisitionOptimizer(space, 'lbfgs', model=model)
    acq = GPyOpt.acquisitions.AcquisitionEI(model, space, aopt, None, 0.01)
    bo = GPyOpt.methods.ModularBayesianOptimization(model, space, objective, acq, GPyOpt.core.evaluators.Sequential(acq), X_init)
    bo.run_optimization(max_iter=iters)
    assert bo.modular_optimization and bo.X.shape == (5 + iters, 2)
    assert_allclose(bo.X, ref.X, rtol=1e-9, atol=1e-9)
    assert_allclose(bo.Y, ref.Y, rtol=1e-9, atol=1e-9)


@pytest.mark.gpu
@pytest.mark.parametrize("acq", ["EI", "LCB"])
def test_bo_branin_trajectory_matches_oracle(acq):
    """BASELINE.json config 1: BO on 2-D Branin, GPRegression RBF, fixed seed -- CUDA numerics vs oracle numerics under the
    identical host loop.  The evaluated points must coincide step by step."""
    iters = 12
    a = _run_bo("cuda", iters, acquisition_type=acq)
    b = _run_bo("oracle", iters, acquisition_type=acq)
    assert a.X.shape == b.X.shape == (5 + iters, 2)
    same = np.all(np.abs(a.X - b.X) <= 1e-5 * (1 + np.abs(b.X)), axis=1)
    first_bad = int(np.argmin(same)) if not same.all() else len(same)
    assert same.all(), "trajectories separate at evaluation %d:\ncuda   %r\noracle %r" % (first_bad, a.X[first_bad], b.X[first_bad])
    assert_allclose(a.Y, b.Y, rtol=1e-4, atol=1e-6)
    assert np.argmin(a.Y) == np.argmin(b.Y)


# ---------------------------------------------------------------------------------------------------------------------
# optimize_restarts with one restart per rank (gloo, world size 2, oracle backend on CPU): same runs, same optimum
# ---------------------------------------------------------------------------------------------------------------------
def _restart_worker(rank, world, port, out):
    import os
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    X, Y = _opt_problem()
    m = make_gpr("oracle", X, Y, GPy.kern.Matern52(3, ARD=True), noise_var=0.1)
    np.random.seed(0)
    m.optimize_restarts(num_restarts=3, optimizer='lbfgs', max_iters=100, verbose=False, distributed=True)
    np.savez(out % rank, x=m.optimizer_array, f=[r.f_opt for r in m.optimization_runs], theta=m[:], rng=np.random.uniform())
    dist.destroy_process_group()


def test_distributed_restarts_match_sequential(tmp_path):
    import os
    import torch.multiprocessing as mp
    out = str(tmp_path / "r%d.npz")
    mp.spawn(_restart_worker, args=(2, 29700 + (os.getpid() % 2000), out), nprocs=2, join=True)
    X, Y = _opt_problem()
    m = make_gpr("oracle", X, Y, GPy.kern.Matern52(3, ARD=True), noise_var=0.1)
    np.random.seed(0)
    m.optimize_restarts(num_restarts=3, optimizer='lbfgs', max_iters=100, verbose=False)
    rng_after = np.random.uniform()
    for r in range(2):
        z = np.load(out % r)
        assert_allclose(z["f"], [run.f_opt for run in m.optimization_runs], rtol=1e-12)
        assert_allclose(z["x"], m.optimizer_array, rtol=1e-12)
        assert_allclose(z["theta"], m[:], rtol=1e-12)
        assert z["rng"] == rng_after            # the global NumPy stream advanced exactly as in the sequential run


@pytest.mark.parametrize("backend", BACKENDS)
def test_concurrent_restarts_match_sequential(backend):
    """optimize_restarts(concurrent=K): K restarts at a time in host threads (own model copy and stream each).  Same runs,
    optimum, parameters and NumPy stream as the sequential loop."""
    X, Y = _opt_problem()
    res = {}
    for conc in (0, 3):
        m = make_gpr(backend, X, Y, GPy.kern.Matern52(3, ARD=True), noise_var=0.1)
        m.Gaussian_noise.constrain_bounded(1e-9, 1e6, warning=False)
        np.random.seed(0)
        m.optimize_restarts(num_restarts=5, optimizer='lbfgs', max_iters=60, verbose=False, concurrent=conc)
        res[conc] = ([r.f_opt for r in m.optimization_runs], [r.funct_eval for r in m.optimization_runs], m.optimizer_array.copy(),
                     m[:].copy(), float(m.log_likelihood()), np.random.uniform())
    assert len(res[3][0]) == 5
    # the device path is bitwise reproducible (fixed-order reductions); the CPU oracle's multi-threaded BLAS is not when several
    # host threads call it at once, so its runs only agree to rounding
    exact = backend == "cuda"
    assert_allclose(res[3][0], res[0][0], rtol=1e-12 if exact else 1e-7)
    if exact:
        assert res[3][1] == res[0][1]
    assert_allclose(res[3][2], res[0][2], rtol=1e-12 if exact else 1e-4, atol=0 if exact else 1e-5)
    assert_allclose(res[3][3], res[0][3], rtol=1e-12 if exact else 1e-4, atol=0 if exact else 1e-6)
    assert_allclose(res[3][4], res[0][4], rtol=1e-12 if exact else 1e-7)
    assert res[3][5] == res[0][5]


def _anchor_bo(distributed):
    np.random.seed(7)
    model = make_gpmodel("oracle", kernel=GPy.kern.Matern52(2), exact_feval=True, verbose=False, optimize_restarts=1)
    bo = GPyOpt.methods.BayesianOptimization(branin, domain=BRANIN_DOMAIN, model=model, acquisition_type='EI', exact_feval=True,
                                             initial_design_numdata=6, evaluator_type='local_penalization', batch_size=2,
                                             distributed_anchors=distributed)
    bo.run_optimization(max_iter=2)
    return bo.X.copy(), bo.Y.copy(), np.random.uniform()


def _anchor_worker(rank, world, port, out):
    import os
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    X, Y, rng = _anchor_bo(True)
    np.savez(out % rank, X=X, Y=Y, rng=rng)
    dist.destroy_process_group()


def test_distributed_anchor_refinement_matches_sequential(tmp_path):
    """AcquisitionOptimizer(distributed_anchors=True): one L-BFGS-B refinement per rank (gloo, world size 2); the BO run
    (local-penalisation batches) proposes the same points as the sequential loop and leaves the NumPy stream in the same state."""
    import os
    import torch.multiprocessing as mp
    out = str(tmp_path / "a%d.npz")
    mp.spawn(_anchor_worker, args=(2, 31700 + (os.getpid() % 2000), out), nprocs=2, join=True)
    X, Y, rng = _anchor_bo(False)
    for r in range(2):
        z = np.load(out % r)
        assert_allclose(z["X"], X, rtol=0, atol=1e-12)
        assert_allclose(z["Y"], Y, rtol=0, atol=1e-12)
        assert z["rng"] == rng


# ---------------------------------------------------------------------------------------------------------------------
# The reference's local Gower mixed-variable kernel patch (stationary.py:116-135), as run.py uses it (Gower=True)
# ---------------------------------------------------------------------------------------------------------------------
GOWER_DOMAIN = [{'name': 'a', 'type': 'discrete', 'domain': (0, 1, 2, 3)}, {'name': 'x', 'type': 'continuous', 'domain': (-5., 10.)},
                {'name': 'b', 'type': 'discrete', 'domain': (1, 2, 3)}, {'name': 'y', 'type': 'continuous', 'domain': (1., 15.)}]


@pytest.mark.parametrize("backend", BACKENDS)
def test_gower_kernel_public_classes(backend, gower_golden):
    g = gower_golden
    space = GPyOpt.Design_space(GOWER_DOMAIN)
    assert space.get_continuous_dims() == list(g["cont_dims"]) and space.get_discrete_dims() == list(g["disc_dims"])
    assert_allclose(space.lengthscales(), g["ranges"])
    w = np.linalg.eigvalsh(g["K"] + (g["noise"] + 1e-8) * np.eye(g["K"].shape[0]))
    ct = max(1.0, w[-1] / w[0] * 2.2e-16 / 1e-12)
    k = GPy.kern.Matern52(4, variance=g["variance"], lengthscale=g["lengthscale"], ARD=g["ard"], Gower=True, space=space)
    if backend == "cuda":      # Kern contract on the device
        assert_allclose(k.K(g["X"]), g["K"], rtol=1e-9)
        assert_allclose(k.K(g["Xs"], g["X"]), g["K_cross"], rtol=1e-9)
        k.update_gradients_full(g["G_sq"], g["X"])
        assert_allclose(np.ravel(k.variance.gradient), g["ugf_sq_var"].ravel(), rtol=1e-7)
        assert_allclose(np.ravel(k.lengthscale.gradient), g["ugf_sq_len"].ravel(), rtol=1e-7)
        k.update_gradients_full(g["G_rect"], g["Xs"], g["X"])
        assert_allclose(np.ravel(k.variance.gradient), g["ugf_rect_var"].ravel(), rtol=1e-7)
        assert_allclose(np.ravel(k.lengthscale.gradient), g["ugf_rect_len"].ravel(), rtol=1e-7)
    m = make_gpr(backend, g["X"], g["Y"], k, noise_var=g["noise"])
    assert_allclose(m.log_likelihood(), g["logL"], rtol=1e-9 * ct)
    assert_allclose(np.ravel(k.variance.gradient), g["grad_var"].ravel(), rtol=1e-7 * ct)
    assert_allclose(np.ravel(k.lengthscale.gradient), g["grad_len"].ravel(), rtol=1e-7 * ct)
    assert_allclose(np.ravel(m.likelihood.variance.gradient), g["grad_noise"].ravel(), rtol=1e-7 * ct)
    assert_allclose(m.posterior.woodbury_chol, g["L"], rtol=1e-9, atol=1e-12)
    mu, var = m.predict(g["Xs"])
    assert_allclose(mu, g["pred_mu"], rtol=1e-9 * ct, atol=1e-11 * ct)
    assert_allclose(var, g["pred_var"], rtol=1e-9 * ct, atol=1e-12 * ct)
    gm = make_gpmodel(backend, exact_feval=False, verbose=False)
    gm.model = m
    mm, ss, dmdx, dsdx = gm.predict_withGradients(g["Xs"])
    assert_allclose(mm, g["gpm_m"], rtol=1e-9 * ct, atol=1e-11 * ct)
    assert_allclose(ss, g["gpm_s"], rtol=1e-8 * ct)
    assert_allclose(dmdx, g["gpm_dmdx"], rtol=1e-7 * ct, atol=1e-9 * ct * np.abs(g["gpm_dmdx"]).max())
    assert_allclose(dsdx, g["gpm_dsdx"], rtol=1e-7 * ct, atol=1e-9 * ct * np.abs(g["gpm_dsdx"]).max())
    assert_allclose(gm.get_fmin(), g["fmin"], rtol=1e-9 * ct)
    unconstrained = GPyOpt.Design_space([{'name': 'x', 'type': 'continuous', 'domain': (-10, 20), 'dimensionality': 4}])
    ei = GPyOpt.acquisitions.AcquisitionEI(gm, unconstrained, optimizer=None, jitter=0.01)
    lcb = GPyOpt.acquisitions.AcquisitionLCB(gm, unconstrained, optimizer=None, exploration_weight=2)
    f, df = ei.acquisition_function_withGradients(g["Xs"])
    assert_allclose(f, g["ei_f"], rtol=1e-7 * ct, atol=1e-14)
    assert_allclose(df, g["ei_df"], rtol=1e-6 * ct, atol=1e-9 * ct * np.abs(g["ei_df"]).max())
    f, df = lcb.acquisition_function_withGradients(g["Xs"])
    assert_allclose(f, g["lcb_f"], rtol=1e-9 * ct, atol=1e-11 * ct)
    assert_allclose(df, g["lcb_df"], rtol=1e-7 * ct, atol=1e-9 * ct * np.abs(g["lcb_df"]).max())
    lp = GPyOpt.acquisitions.AcquisitionLP(gm, unconstrained, None, ei)
    lp.update_batches(g["lp_Xb"], 1.5, float(g["Y"].min()))
    assert_allclose(lp.acquisition_function(g["Xs"][8:]), g["lp_f"], rtol=1e-7 * ct, atol=1e-7 * ct)


@pytest.mark.parametrize("backend", BACKENDS)
def test_bo_mixed_space_gower_local_penalization(backend):
    """The configuration of run.py:1207-1224 in miniature: mixed discrete / continuous space, Gower=True, exact_feval,
    local-penalisation batches, then scoring an explicit candidate list with the (penalised) acquisition (run.py:1234-1253)."""
    def f(X):
        X = np.atleast_2d(X)
        return (np.sin(X[:, 1] / 3) + 0.3 * X[:, 0] - 0.2 * (X[:, 2] == 2) + 0.01 * X[:, 3] ** 2).reshape(-1, 1)
    np.random.seed(7)
    space = GPyOpt.Design_space(GOWER_DOMAIN)
    X0 = GPyOpt.experiment_design.initial_design('random', space, 12)
    kw = dict(f=None, domain=GOWER_DOMAIN, X=X0, Y=f(X0), acquisition_type='EI', normalize_Y=True, exact_feval=True,
              evaluator_type='local_penalization', batch_size=3, de_duplication=True, Gower=True, noise_var=0, optimize_restarts=2)
    if backend == "oracle":
        kw["model"] = OB.OracleGPModel(noise_var=0, exact_feval=True, optimize_restarts=2, verbose=False, Gower=True, space=space)
    bo = GPyOpt.methods.BayesianOptimization(**kw)
    Xn = bo.suggest_next_locations()
    assert Xn.shape == (3, 4)
    assert set(Xn[:, 0]) <= {0., 1., 2., 3.} and set(Xn[:, 2]) <= {1., 2., 3.}          # discrete variables are rounded to their domain
    assert bo.model.model.kern.gower_config() is not None
    cand = GPyOpt.experiment_design.initial_design('random', space, 200)
    acq = bo.evaluator.acquisition
    acq.update_batches(None, None, None)
    v0 = acq.acquisition_function(cand)
    assert v0.shape == (200,) and np.all(np.isfinite(v0))
    L = GPyOpt.core.evaluators.batch_local_penalization.estimate_L(bo.model.model, space.get_bounds())
    acq.update_batches(cand[int(np.argmin(v0))], L, bo.model.model.Y.min())
    v1 = acq.acquisition_function(cand)
    assert np.all(np.isfinite(v1)) and np.all(v1 >= v0 - 1e-12)    # the penalisers only add -log Phi(.) >= 0 to the minimised value
    assert acq.r_x0.shape == (1,) and acq.s_x0.shape == (1,)


# ---------------------------------------------------------------------------------------------------------------------
# check_kernel_gradient_functions (GPy/GPy/testing/kernel_tests.py:23-349; instantiated for Matern52 :414-417 and RBF(ARD)
# :419-422 with N = 10, N2 = 20, D = 5 :352-356): the Kern contract against central differences of sum(dL_dK * K)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("kname,ard", [("Matern52", False), ("RBF", True), ("Matern52", True), ("RBF", False)])
def test_kernel_gradient_functions(kname, ard):
    rs = np.random.RandomState(31)
    N, N2, D = 10, 20, 5
    X, X2 = rs.randn(N, D), rs.randn(N2, D)
    k = getattr(GPy.kern, kname)(D, ARD=ard)
    k.variance[...] = 1.0 + 0.1 * rs.randn()                      # kernel.randomize(loc=1, scale=0.1)
    k.lengthscale[...] = 1.0 + 0.1 * rs.randn(k.lengthscale.size)

    # positive semi-definiteness and Kdiag (kernel_tests.py:47-53, Kern_check_dKdiag_dtheta)
    Kxx = k.K(X)
    assert np.linalg.eigvalsh(Kxx).min() > -1e-10
    assert_allclose(k.Kdiag(X), np.diag(Kxx), rtol=1e-12)
    assert_allclose(Kxx, Kxx.T, rtol=0, atol=0)

    for Xb, G in ((None, rs.rand(N, N)), (X2, rs.rand(N, N2))):
        f = lambda: float(np.sum(G * k.K(X, Xb)))                 # noqa: E731  Kern_check_model.log_likelihood
        # dK/dtheta (Kern_check_dK_dtheta)
        k.update_gradients_full(G, X, Xb)
        g_var, g_len = float(np.ravel(k.variance.gradient)[0]), np.ravel(k.lengthscale.gradient).copy()
        h = 1e-6
        v0 = float(k.variance.values[0])
        k.variance[...] = v0 + h
        fp = f()
        k.variance[...] = v0 - h
        fm = f()
        k.variance[...] = v0
        assert_allclose(g_var, (fp - fm) / (2 * h), rtol=1e-6)
        l0 = k.lengthscale.values.copy()
        for q in range(l0.size):
            lp, lm = l0.copy(), l0.copy()
            lp[q] += h
            lm[q] -= h
            k.lengthscale[...] = lp
            fp = f()
            k.lengthscale[...] = lm
            fm = f()
            k.lengthscale[...] = l0
            assert_allclose(g_len[q], (fp - fm) / (2 * h), rtol=1e-5, atol=1e-8)
        # dK/dX (Kern_check_dK_dX): gradient of sum(G * K(X, Xb)) with respect to X
        gX = k.gradients_X(G, X, Xb)
        assert gX.shape == X.shape
        num = np.zeros_like(X)
        for i in range(N):
            for q in range(D):
                Xp, Xm = X.copy(), X.copy()
                Xp[i, q] += h
                Xm[i, q] -= h
                if Xb is None:
                    num[i, q] = (np.sum(G * k.K(Xp)) - np.sum(G * k.K(Xm))) / (2 * h)
                else:
                    num[i, q] = (np.sum(G * k.K(Xp, Xb)) - np.sum(G * k.K(Xm, Xb))) / (2 * h)
        assert_allclose(gX, num, rtol=1e-5, atol=1e-7)
    assert np.all(k.gradients_X_diag(np.ones(N), X) == 0.0)      # stationary.py:366-367


# ---------------------------------------------------------------------------------------------------------------------
# set_XY that only appends rows under unchanged hyper-parameters extends the resident factorisation
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_set_xy_append_is_incremental_and_equal_to_a_rebuild():
    rs = np.random.RandomState(11)
    D, n0 = 3, 120
    X = rs.uniform(0, 1, (n0 + 40, D))
    Y = np.sin(X @ np.array([3., 2., 1.]))[:, None] + 0.05 * rs.randn(X.shape[0], 1)
    m = GPy.models.GPRegression(X[:n0], Y[:n0], kernel=GPy.kern.Matern52(D, ARD=True), noise_var=0.05)
    inf = m.inference_method
    assert inf.incremental and inf.n_appends == 0
    Xc = rs.uniform(0, 1, (30, D))
    n = n0
    for b in (1, 5, 9, 1):                        # crosses the 128 boundary at the third step
        n += b
        Yn = (Y[:n] - Y[:n].mean()) / Y[:n].std()  # GPyOpt re-normalises every target on every step (bo.py:246-247)
        m.set_XY(X[:n], Yn)
        ref = GPy.models.GPRegression(X[:n], Yn, kernel=GPy.kern.Matern52(D, ARD=True), noise_var=0.05)
        assert ref.inference_method.n_appends == 0
        assert_allclose(m.log_likelihood(), ref.log_likelihood(), rtol=1e-10)
        assert_allclose(m.gradient, ref.gradient, rtol=1e-7, atol=1e-9)
        for a, r in zip(m.predict(Xc), ref.predict(Xc)):
            assert_allclose(a, r, rtol=1e-8, atol=1e-11)
        for a, r in zip(m.predictive_gradients(Xc), ref.predictive_gradients(Xc)):
            assert_allclose(a, r, rtol=1e-7, atol=1e-9 * np.abs(r).max())
        assert_allclose(m.posterior.woodbury_inv, ref.posterior.woodbury_inv, rtol=1e-7, atol=1e-9 * np.abs(ref.posterior.woodbury_inv).max())
    assert inf.n_appends == 4
    # anything but "same rows + more rows, same hyper-parameters" takes the full path
    Xp = X[:n + 3].copy()
    Xp[0, 0] += 1e-3
    m.set_XY(Xp, Y[:n + 3])
    assert inf.n_appends == 4
    m.set_XY(X[:n + 3], Y[:n + 3])                 # first row differs from the resident one again
    assert inf.n_appends == 4
    m.kern.lengthscale[:] = 0.7
    m.set_XY(X[:n + 6], Y[:n + 6])
    assert inf.n_appends == 5                      # the hyper-parameter change refitted first; the set_XY then appended
    m.set_XY(X[:n + 2], Y[:n + 2])                 # fewer rows
    assert inf.n_appends == 5
    ref = GPy.models.GPRegression(X[:n + 2], Y[:n + 2], kernel=GPy.kern.Matern52(D, ARD=True, lengthscale=0.7), noise_var=0.05)
    assert_allclose(m.log_likelihood(), ref.log_likelihood(), rtol=1e-10)


@pytest.mark.gpu
def test_bo_frozen_hyperparameters_incremental_matches_rebuild(monkeypatch):
    """GPModel(max_iters=0): the hyper-parameters never move, so every updateModel is an append.  Same trajectory as rebuilding
    the model on every step (what the reference does, gpmodel.py:78-93)."""
    from gaussian_process_optimization_b200 import models as _models
    res = {}
    for incremental in (True, False):
        monkeypatch.setattr(_models.ExactGaussianInference, "INCREMENTAL", incremental)
        np.random.seed(3)
        model = GPyOpt.models.GPModel(kernel=GPy.kern.Matern52(2, ARD=True, lengthscale=[4., 5.], variance=2.), exact_feval=True,
                                      verbose=False, max_iters=0)
        bo = GPyOpt.methods.BayesianOptimization(branin, domain=BRANIN_DOMAIN, model=model, acquisition_type='EI',
                                                 exact_feval=True, initial_design_numdata=122, initial_design_type='random')
        bo.run_optimization(max_iter=10)
        res[incremental] = (bo.X.copy(), bo.Y.copy(), model.model.inference_method.n_appends)
    assert res[True][2] >= 9 and res[False][2] == 0
    assert res[True][0].shape == res[False][0].shape
    assert_allclose(res[True][0], res[False][0], rtol=0, atol=1e-5 * 15)


@pytest.mark.gpu
def test_set_xy_append_under_the_gower_patch():
    """The appended block rows are built with the same (Gower) covariance as a rebuild (run.py's configuration: Gower=True)."""
    space = GPyOpt.Design_space(GOWER_DOMAIN)
    np.random.seed(2)
    X = GPyOpt.experiment_design.initial_design('random', space, 150)
    Y = (np.sin(X[:, 1] / 3) + 0.3 * X[:, 0] - 0.2 * (X[:, 2] == 2) + 0.01 * X[:, 3] ** 2).reshape(-1, 1)
    mk = lambda: GPy.kern.Matern52(4, variance=1.3, ARD=False, Gower=True, space=space)   # noqa: E731
    m = GPy.models.GPRegression(X[:120], Y[:120], kernel=mk(), noise_var=0.05)
    for n in (125, 131, 150):
        m.set_XY(X[:n], Y[:n])
        ref = GPy.models.GPRegression(X[:n], Y[:n], kernel=mk(), noise_var=0.05)
        assert_allclose(m.log_likelihood(), ref.log_likelihood(), rtol=1e-10)
        assert_allclose(m.gradient, ref.gradient, rtol=1e-7, atol=1e-9)
        for a, r in zip(m.predict(X[140:150]), ref.predict(X[140:150])):
            assert_allclose(a, r, rtol=1e-8, atol=1e-11)
    assert m.inference_method.n_appends == 3
