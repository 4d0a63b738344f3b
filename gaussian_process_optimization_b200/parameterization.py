"""Minimal parameter machinery with the paramz behaviours GPy/GPyOpt rely on for the exact-GP path.

paramz is an un-vendored dependency of the reference (GPy/setup.py:162) and is NOT under /root/reference; what is restated
here follows its published behaviour and is covered by our own tests (parity of the optimiser
trajectory is therefore unpinned against paramz itself):

  * Param: an ndarray with a name, a `.gradient`, and a constraint (Logexp positive / Logistic bounded / fixed);
    every write notifies the owning model, which re-runs `parameters_changed` (observer pattern), unless updates are off.
  * Parameterized: ordered tree of parameters (link order = order in `m[:]`, stationary.py:83, core/gp.py:108-109).
  * Model: optimizer_array <-> param_array through the transforms, objective / gradients with the chain rule
    (`gradfactor`), `optimize` (L-BFGS-B through scipy.optimize.fmin_l_bfgs_b, maxfun = maxiter = max_iters, SciPy defaults
    otherwise), `optimize_restarts`, `randomize`, `checkgrad`.
"""
import re

import numpy as np

_LIM_VAL = 36.0
_LOG_LIM_VAL = np.log(np.finfo(np.float64).max)


class Logexp(object):
    """theta = log(1 + exp(x)) (positive)."""
    domain = "positive"

    def f(self, x):
        x = np.asarray(x, dtype=np.float64)
        return np.where(x > _LIM_VAL, x, np.log1p(np.exp(np.clip(x, -_LOG_LIM_VAL, _LIM_VAL))))

    def finv(self, f):
        f = np.asarray(f, dtype=np.float64)
        return np.where(f > _LIM_VAL, f, np.log(np.expm1(f)))

    def gradfactor(self, f, df):
        f = np.asarray(f, dtype=np.float64)
        return df * np.where(f > _LIM_VAL, 1., -np.expm1(-f))

    def __str__(self):
        return "+ve"


class Logistic(object):
    """theta = lower + (upper - lower) / (1 + exp(-x)) (bounded)."""
    domain = "bounded"

    def __init__(self, lower, upper):
        assert lower < upper
        self.lower, self.upper = float(lower), float(upper)
        self.difference = self.upper - self.lower

    def f(self, x):
        x = np.array(x, dtype=np.float64)
        x[x < -300.] = -300.
        return self.lower + self.difference / (1. + np.exp(-x))

    def finv(self, f):
        f = np.asarray(f, dtype=np.float64)
        return np.log(np.clip(f - self.lower, 1e-10, np.inf) / np.clip(self.upper - f, 1e-10, np.inf))

    def gradfactor(self, f, df):
        f = np.asarray(f, dtype=np.float64)
        return df * (f - self.lower) * (self.upper - f) / self.difference

    def __str__(self):
        return "{},{}".format(self.lower, self.upper)


class Param(np.ndarray):
    """A named array of parameter values.  Arithmetic yields plain ndarrays; writes notify the owning model."""

    def __new__(cls, name, value, default_constraint=None):
        obj = np.atleast_1d(np.array(value, dtype=np.float64)).view(cls)
        obj.name = name
        obj.gradient = np.zeros(obj.shape)
        obj._constraint = default_constraint
        obj._fixed = False
        obj._parent = None
        return obj

    def __array_finalize__(self, obj):
        if obj is None:
            return
        self.name = getattr(obj, "name", None)
        self.gradient = getattr(obj, "gradient", None)
        self._constraint = getattr(obj, "_constraint", None)
        self._fixed = getattr(obj, "_fixed", False)
        self._parent = None

    def __array_wrap__(self, out, context=None, return_scalar=False):
        out = np.asarray(out)
        return out[()] if return_scalar else out

    def __reduce__(self):
        return (_rebuild_param, (self.name, np.asarray(self).copy(), self._constraint, self._fixed))

    @property
    def values(self):
        return self.view(np.ndarray)

    def __setitem__(self, key, value):
        np.ndarray.__setitem__(self, key, value)
        self._notify()

    def _notify(self):
        if self._parent is not None:
            self._parent._child_changed()

    # -- constraints (paramz Constrainable) ----------------------------------------------------------------------------
    def constrain_fixed(self, value=None, warning=True, trigger_parent=True):
        if value is not None:
            np.ndarray.__setitem__(self, Ellipsis, value)
        self._fixed = True
        self._notify()
        return self
    fix = constrain_fixed

    def unconstrain_fixed(self):
        self._fixed = False
    unfix = unconstrain_fixed

    def constrain(self, transform, warning=True):
        self._constraint = transform
        self._fixed = False
        if transform is not None:  # move infeasible values into the domain like paramz's initialize()
            v = self.values
            if isinstance(transform, Logexp):
                bad = v <= 0
                if bad.any():
                    np.ndarray.__setitem__(self, bad, np.abs(v[bad]) + 1e-300)
            elif isinstance(transform, Logistic):
                bad = (v <= transform.lower) | (v >= transform.upper)
                if bad.any():
                    np.ndarray.__setitem__(self, bad, transform.f(np.zeros(int(bad.sum()))))
        self._notify()
        return self

    def constrain_positive(self, warning=True):
        return self.constrain(Logexp(), warning)

    def constrain_bounded(self, lower, upper, warning=True):
        return self.constrain(Logistic(lower, upper), warning)

    def unconstrain(self):
        self._constraint = None
        self._fixed = False

    def copy(self):
        p = Param(self.name, np.asarray(self).copy(), self._constraint)
        p._fixed = self._fixed
        p.gradient = np.array(self.gradient, copy=True)
        return p

    def __str__(self):
        return "{}: {}".format(self.name, np.asarray(self))


def _rebuild_param(name, value, constraint, fixed):
    p = Param(name, value, constraint)
    p._fixed = fixed
    return p


class Parameterized(object):
    """Ordered container of Params / Parameterized children with change notification to the root."""

    def __init__(self, name=None, *a, **kw):
        object.__setattr__(self, "_in_init_", True)
        self.name = name or self.__class__.__name__
        self.parameters = []
        self._parent = None
        self._updates = True
        object.__setattr__(self, "_in_init_", False)

    # -- tree ------------------------------------------------------------------------------------------------------------
    def link_parameter(self, param, index=None):
        param._parent = self
        if index is None:
            self.parameters.append(param)
        else:
            self.parameters.insert(index, param)
        pname = re.sub(r"\W", "_", param.name)
        object.__setattr__(self, pname, param)

    def link_parameters(self, *params):
        for p in params:
            self.link_parameter(p)

    def unlink_parameter(self, param):
        self.parameters = [p for p in self.parameters if p is not param]
        param._parent = None

    def __setattr__(self, key, value):
        cur = self.__dict__.get(key, None)
        if isinstance(cur, Param) and not isinstance(value, Param):
            cur[...] = value  # in-place write -> notification
            return
        object.__setattr__(self, key, value)

    def flattened_parameters(self):
        out = []
        for p in self.parameters:
            if isinstance(p, Param):
                out.append(p)
            else:
                out.extend(p.flattened_parameters())
        return out

    def hierarchy_name(self):
        if self._parent is None:
            return re.sub(r"\W", "_", self.name)
        return self._parent.hierarchy_name() + "." + re.sub(r"\W", "_", self.name)

    # -- notification ------------------------------------------------------------------------------------------------
    def _root(self):
        r = self
        while r._parent is not None:
            r = r._parent
        return r

    def _child_changed(self):
        root = self._root()
        if getattr(root, "_in_init_", False):
            return
        if root._updates:
            root._trigger()
        else:
            root._dirty = True

    def _trigger(self):
        for p in self.parameters:
            if isinstance(p, Parameterized):
                p.parameters_changed()
        self.parameters_changed()

    def parameters_changed(self):
        pass

    def update_model(self, updates=None):
        """paramz Updateable.update_model: query (None) or switch; switching on triggers an update."""
        root = self._root()
        if updates is None:
            return root._updates
        root._updates = bool(updates)
        if updates:
            root._trigger()

    # -- flat views ----------------------------------------------------------------------------------------------------
    @property
    def size(self):
        return int(sum(p.size for p in self.flattened_parameters()))

    @property
    def param_array(self):
        ps = self.flattened_parameters()
        return np.concatenate([p.values.ravel() for p in ps]) if ps else np.zeros(0)

    @param_array.setter
    def param_array(self, x):
        self._set_params(np.asarray(x, dtype=np.float64).ravel())
        self._child_changed()

    def _set_params(self, x):
        i = 0
        for p in self.flattened_parameters():
            np.ndarray.__setitem__(p, Ellipsis, x[i:i + p.size].reshape(p.shape))
            i += p.size

    @property
    def gradient(self):
        ps = self.flattened_parameters()
        return np.concatenate([np.asarray(p.gradient, dtype=np.float64).ravel() * np.ones(p.size) for p in ps]) if ps else np.zeros(0)

    def parameter_names_flat(self, include_fixed=False):
        names = []
        for p in self.flattened_parameters():
            if p._fixed and not include_fixed:
                continue
            base = (p._parent.hierarchy_name() + "." if p._parent is not None else "") + p.name
            if p.size == 1:
                names.append(base)
            else:
                names.extend("{}[[{}]]".format(base, i) for i in range(p.size))
        return np.array(names)

    def __getitem__(self, key):
        if isinstance(key, str):
            rx = re.compile(key)
            vals = [p.values.ravel() for p in self.flattened_parameters()
                    if rx.search((p._parent.hierarchy_name() + "." if p._parent is not None else "") + p.name)]
            return np.concatenate(vals) if vals else np.zeros(0)
        return self.param_array[key]

    def __setitem__(self, key, value):
        if isinstance(key, str):
            rx = re.compile(key)
            for p in self.flattened_parameters():
                if rx.search((p._parent.hierarchy_name() + "." if p._parent is not None else "") + p.name):
                    np.ndarray.__setitem__(p, Ellipsis, value)
            self._child_changed()
            return
        x = self.param_array
        x[key] = value
        self.param_array = x

    # -- optimiser view --------------------------------------------------------------------------------------------------
    def _free(self):
        return [p for p in self.flattened_parameters() if not p._fixed]

    def _size_transformed(self):
        return int(sum(p.size for p in self._free()))

    @property
    def is_fixed(self):
        return self._size_transformed() == 0

    @property
    def optimizer_array(self):
        out = []
        for p in self._free():
            v = p.values.ravel()
            out.append(p._constraint.finv(v) if p._constraint is not None else v.copy())
        return np.concatenate(out) if out else np.zeros(0)

    @optimizer_array.setter
    def optimizer_array(self, x):
        x = np.asarray(x, dtype=np.float64).ravel()
        i = 0
        for p in self._free():
            xi = x[i:i + p.size]
            v = p._constraint.f(xi) if p._constraint is not None else xi
            np.ndarray.__setitem__(p, Ellipsis, np.asarray(v, dtype=np.float64).reshape(p.shape))
            i += p.size
        self._child_changed()

    def _transform_gradients(self, g):
        """Chain rule through the constraints, free parameters only (paramz Parameterized._transform_gradients)."""
        g = np.asarray(g, dtype=np.float64).ravel()
        out, i = [], 0
        for p in self.flattened_parameters():
            gi = g[i:i + p.size]
            i += p.size
            if p._fixed:
                continue
            out.append(p._constraint.gradfactor(p.values.ravel(), gi) if p._constraint is not None else gi.copy())
        return np.concatenate(out) if out else np.zeros(0)

    def copy(self):
        import copy as _copy
        c = _copy.deepcopy(self)
        c._parent = None
        return c


class ObjectiveRun(object):
    """What paramz keeps per optimisation run (optimization_runs[i]): x_opt, f_opt, funct_eval, status."""

    def __init__(self, x_opt, f_opt, funct_eval, status):
        self.x_opt, self.f_opt, self.funct_eval, self.status = x_opt, f_opt, funct_eval, status


class Model(Parameterized):
    """paramz.Model restated: objective = -log_likelihood - log_prior (GPy/GPy/core/model.py:96-127; no priors here)."""

    _allowed_failures = 10

    def __init__(self, name):
        super(Model, self).__init__(name)
        self.optimization_runs = []
        self._fail_count = 0
        self.preferred_optimizer = "lbfgsb"
        self.obj_grads = None

    def log_likelihood(self):
        raise NotImplementedError

    def _log_likelihood_gradients(self):
        return self.gradient

    def objective_function(self):
        return -float(self.log_likelihood())

    def objective_function_gradients(self):
        return -self._log_likelihood_gradients()

    def _objective_grads(self, x):
        try:
            self.optimizer_array = x
            obj_f, self.obj_grads = self.objective_function(), self._transform_gradients(self.objective_function_gradients())
            self._fail_count = 0
        except (np.linalg.LinAlgError, ZeroDivisionError, ValueError):
            if self._fail_count >= self._allowed_failures:
                raise
            self._fail_count += 1
            obj_f = np.finfo(np.float64).max
            self.obj_grads = np.clip(self._transform_gradients(self.objective_function_gradients()), -1e10, 1e10)
        return obj_f, self.obj_grads

    def optimize(self, optimizer=None, start=None, messages=False, max_iters=1000, ipython_notebook=True,
                 clear_after_finish=False, **kwargs):
        """paramz Model.optimize with the L-BFGS-B wrapper ('lbfgs', 'bfgs', 'lbfgsb' all resolve to it in paramz)."""
        if self.is_fixed or self.size == 0:
            print("nothing to optimize")
            return
        if not self.update_model():
            print("updates were off, setting updates on again")
            self.update_model(True)
        if start is None:
            start = self.optimizer_array
        name = (optimizer or self.preferred_optimizer).lower()
        if "bfgs" not in name:
            raise NotImplementedError("only the L-BFGS-B optimiser of the reference path is provided (got %r)" % optimizer)
        import scipy.optimize
        opt_dict = {}
        if kwargs.get("gtol") is not None:
            opt_dict["pgtol"] = kwargs["gtol"]
        if kwargs.get("bfgs_factor") is not None:
            opt_dict["factr"] = kwargs["bfgs_factor"]
        res = scipy.optimize.fmin_l_bfgs_b(self._objective_grads, start, maxfun=max_iters, maxiter=max_iters, **opt_dict)
        x_opt = res[0]
        f_opt = self._objective_grads(x_opt)[0]
        rc = ["Converged", "Maximum number of f evaluations reached", "Error"]
        status = rc[res[2]["warnflag"]]
        if res[2]["warnflag"] == 2:
            status = "Error" + str(res[2]["task"])
        self.optimizer_array = x_opt
        run = ObjectiveRun(x_opt, f_opt, res[2]["funcalls"], status)
        self.optimization_runs.append(run)
        return run

    def _optimize_restarts_distributed(self, num_restarts, robust, verbose, group, **kwargs):
        """One restart per rank: the restarts are independent once their starting points are drawn, so rank r
        runs restarts r, r + world, ... and ONE all-reduce of (f_opt, x_opt) per restart row makes every rank adopt the same
        best run.  Every rank must hold the same model and the same np.random state: the starts are drawn exactly as the
        sequential loop draws them (one normal(size=n_free) vector per restart i > 0, GPy/GPy/core/__init__.py:19-43), which
        keeps the global RNG stream -- and therefore the BO trajectory -- identical to the sequential run."""
        import torch
        import torch.distributed as dist
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        n = self._size_transformed()
        starts = [self.optimizer_array.copy()] + [np.random.normal(size=n) for _ in range(1, num_restarts)]
        table = np.full((num_restarts, 3 + n), np.inf)        # f_opt, funct_eval, ok flag, x_opt
        for i in range(rank, num_restarts, world):
            try:
                if i > 0:                                     # restart 0 continues from the current parameters (sequential loop)
                    self.optimizer_array = starts[i]
                run = self.optimize(**kwargs)                 # reads its start back through the parameter transforms, like randomize() + optimize()
                table[i, 0], table[i, 1], table[i, 2], table[i, 3:] = run.f_opt, run.funct_eval, 1.0, run.x_opt
                self.optimization_runs.pop()                  # re-appended below in restart order on every rank
            except Exception as e:
                if not robust:
                    raise e
                print("Warning - optimization restart {0}/{1} failed".format(i + 1, num_restarts))
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
        t = torch.from_numpy(table).to(dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)  # every row was written by exactly one rank, +inf elsewhere
        table = t.cpu().numpy()
        done = [i for i in range(num_restarts) if np.isfinite(table[i, 2])]
        for i in done:
            self.optimization_runs.append(ObjectiveRun(table[i, 3:].copy(), float(table[i, 0]), int(table[i, 1]), "distributed"))
            if verbose:
                print("Optimization restart {0}/{1}, f = {2}".format(i + 1, num_restarts, table[i, 0]))
        return done, table

    def _optimize_restarts_concurrent(self, num_restarts, robust, verbose, workers, **kwargs):
        """Restarts on one GPU, `workers` at a time: below a few thousand training points a single NLL+grad evaluation is a chain
        of short kernels that leaves most of the 148 SMs idle, and the restarts are independent once their starting points are
        drawn.  Every worker thread owns a copy of the model (its own device workspace and CUDA stream; the C library holds no
        shared mutable state besides a launch counter) and runs restarts w, w + workers, ...; ctypes releases the GIL during
        the device calls, so the kernels of different restarts overlap.  The starts are drawn exactly as the sequential loop
        draws them and every run is deterministic, so runs, optimum and the NumPy stream equal the sequential ones."""
        import threading
        n = self._size_transformed()
        starts = [self.optimizer_array.copy()] + [np.random.normal(size=n) for _ in range(1, num_restarts)]
        runs, errors = [None] * num_restarts, [None] * num_restarts
        dev = None
        try:
            import torch
            if torch.cuda.is_available() and torch.cuda.is_initialized():
                dev = torch.cuda.current_device()
        except ImportError:
            torch = None

        def work(w):
            import contextlib
            ctx = contextlib.nullcontext()
            if dev is not None:
                torch.cuda.set_device(dev)                     # the current device is per host thread
                ctx = torch.cuda.stream(torch.cuda.Stream())
            with ctx:
                m = None
                for i in range(w, num_restarts, workers):
                    try:
                        if m is None:
                            m = self.copy()
                        # like the sequential loop: restart 0 continues from the current parameters, the others from
                        # randomize(); optimize() then reads its start back through the parameter transforms
                        if i > 0:
                            m.optimizer_array = starts[i]
                        runs[i] = m.optimize(**kwargs)
                    except Exception as e:                      # reported in restart order by the caller
                        errors[i] = e
                nat = getattr(getattr(m, "inference_method", None), "_nat", None)
                if nat is not None:
                    nat.close()

        threads = [threading.Thread(target=work, args=(w,)) for w in range(min(workers, num_restarts))]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        done = []
        for i in range(num_restarts):
            if errors[i] is not None:
                if not robust:
                    raise errors[i]
                print("Warning - optimization restart {0}/{1} failed".format(i + 1, num_restarts))
            elif runs[i] is not None:
                self.optimization_runs.append(runs[i])
                done.append(i)
                if verbose:
                    print("Optimization restart {0}/{1}, f = {2}".format(i + 1, num_restarts, runs[i].f_opt))
        return done, runs

    def optimize_restarts(self, num_restarts=10, robust=False, verbose=True, parallel=False, num_processes=None, distributed=False,
                          group=None, concurrent=0, **kwargs):
        if concurrent and concurrent > 1 and num_restarts > 1:
            initial_parameters = self.optimizer_array.copy()
            done, runs = self._optimize_restarts_concurrent(num_restarts, robust, verbose, int(concurrent), **kwargs)
            if done:
                best = done[int(np.argmin([runs[i].f_opt for i in done]))]
                self.optimizer_array = runs[best].x_opt
            else:
                self.optimizer_array = initial_parameters
            return self.optimization_runs
        if distributed:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
                initial_parameters = self.optimizer_array.copy()
                done, table = self._optimize_restarts_distributed(num_restarts, robust, verbose, group, **kwargs)
                if done:
                    best = done[int(np.argmin([table[i, 0] for i in done]))]
                    self.optimizer_array = table[best, 3:]
                else:
                    self.optimizer_array = initial_parameters
                return self.optimization_runs
        initial_length = len(self.optimization_runs)
        initial_parameters = self.optimizer_array.copy()
        for i in range(num_restarts):
            try:
                if i > 0:
                    self.randomize()
                self.optimize(**kwargs)
                if verbose:
                    print("Optimization restart {0}/{1}, f = {2}".format(i + 1, num_restarts, self.optimization_runs[-1].f_opt))
            except Exception as e:
                if robust:
                    print("Warning - optimization restart {0}/{1} failed".format(i + 1, num_restarts))
                else:
                    raise e
        if len(self.optimization_runs) > initial_length:
            i = int(np.argmin([o.f_opt for o in self.optimization_runs[initial_length:]]))
            self.optimizer_array = self.optimization_runs[initial_length + i].x_opt
        else:
            self.optimizer_array = initial_parameters
        return self.optimization_runs

    def randomize(self, rand_gen=None, *args, **kwargs):
        """GPy/GPy/core/__init__.py:19-43: ONE draw of normal(size=n_free) assigned to the optimiser array (no priors)."""
        if rand_gen is None:
            rand_gen = np.random.normal
        x = rand_gen(size=self._size_transformed(), *args, **kwargs)
        updates = self.update_model()
        self._root()._updates = False
        self.optimizer_array = x
        self._root()._updates = updates
        if updates:
            self._trigger()

    def checkgrad(self, verbose=False, step=1e-6, tolerance=1e-3):
        """Central-difference check of the transformed objective gradient (paramz Model.checkgrad, global ratio test)."""
        x = self.optimizer_array.copy()
        dx = np.random.uniform(-1, 1, x.size) * step if x.size else x
        dx = np.where(dx == 0., step, dx)
        f1 = self._objective_grads(x + dx)[0]
        f2 = self._objective_grads(x - dx)[0]
        g = self._objective_grads(x)[1]
        num = (f1 - f2) / 2.0
        ana = float(np.dot(g, dx))
        self.optimizer_array = x
        denom = ana if ana != 0 else 1e-300
        ratio = num / denom
        return bool(abs(1. - ratio) < tolerance or abs(num - ana) < tolerance)
