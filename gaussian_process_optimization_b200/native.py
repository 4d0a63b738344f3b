"""Thin Python wrappers over the C ABI (include/gpb200.h): NumPy arrays in/out (host path, copies inside the call) or CUDA
torch tensors (device path, in place).  No numerics happen here -- every function forwards to libgpb200.so."""
import ctypes

import numpy as np

from . import _lib
from ._lib import KIND_IDS, ACQ_IDS, as_host, check, dptr, is_torch, ptr


def _kind(kind):
    return kind if isinstance(kind, int) else KIND_IDS[kind]


def _ls(lengthscale):
    return np.ascontiguousarray(np.atleast_1d(np.asarray(lengthscale, dtype=np.float64)))


def kern_K(kind, X, X2, variance, lengthscale):
    """Stationary.K(X, X2) (GPy/GPy/kern/src/stationary.py:107-140)."""
    lib = _lib.require_gpu()
    X = as_host(X)
    n, d = X.shape
    ls = _ls(lengthscale)
    if X2 is None:
        m, X2p = 0, None
        out = np.empty((n, n))
    else:
        X2 = as_host(X2)
        m, X2p = X2.shape[0], ptr(X2)
        out = np.empty((n, m))
    check(lib.gpb_kern_K(_kind(kind), d, n, ptr(X), m, X2p, float(variance), dptr(ls), ls.size, ptr(out), out.shape[1], 0,
                         _lib.current_stream()), "kern_K")
    return out


def kern_update_gradients_full(kind, dL_dK, X, X2, variance, lengthscale):
    """Stationary.update_gradients_full (stationary.py:218-238) -> (dvariance, dlengthscale[nls])."""
    lib = _lib.require_gpu()
    X = as_host(X)
    n, d = X.shape
    ls = _ls(lengthscale)
    G = as_host(dL_dK)
    if X2 is None:
        m, X2p = 0, None
        assert G.shape == (n, n)
    else:
        X2 = as_host(X2)
        m, X2p = X2.shape[0], ptr(X2)
        assert G.shape == (n, m)
    out = np.empty(1 + ls.size)
    check(lib.gpb_kern_update_gradients_full(_kind(kind), d, n, ptr(X), m, X2p, ptr(G), G.shape[1], float(variance), dptr(ls), ls.size,
                                             dptr(out), 0, _lib.current_stream()), "update_gradients_full")
    return out[0], out[1:].copy()


def _gower_args(gower, d):
    """gower = (continuous dims, discrete dims, ranges of the continuous dims in that order) -> (int flags[d], ranges[d])."""
    cont, disc, ranges = gower
    flags = np.zeros(d, dtype=np.int32)
    rng = np.ones(d)
    flags[list(disc)] = 1
    for i, q in enumerate(cont):
        rng[q] = float(ranges[i])
    assert len(cont) + len(disc) == d, "every input dimension must be continuous or discrete"
    return flags, rng


def kern_K_gower(kind, X, X2, variance, gower):
    """Stationary.K under the Gower patch (stationary.py:116-135)."""
    lib = _lib.require_gpu()
    X = as_host(X)
    n, d = X.shape
    flags, rng = _gower_args(gower, d)
    if X2 is None:
        m, X2p, out = 0, None, np.empty((n, n))
    else:
        X2 = as_host(X2)
        m, X2p, out = X2.shape[0], ptr(X2), np.empty((n, X2.shape[0]))
    check(lib.gpb_kern_K_gower(_kind(kind), d, n, ptr(X), m, X2p, float(variance), flags.ctypes.data_as(_lib.c_int_p), dptr(rng),
                               ptr(out), out.shape[1], 0, _lib.current_stream()), "kern_K_gower")
    return out


def kern_update_gradients_full_gower(kind, dL_dK, X, X2, variance, lengthscale, gower):
    """Stationary.update_gradients_full under the Gower patch: patched K in the variance term only (stationary.py:224)."""
    lib = _lib.require_gpu()
    X = as_host(X)
    n, d = X.shape
    ls = _ls(lengthscale)
    G = as_host(dL_dK)
    flags, rng = _gower_args(gower, d)
    if X2 is None:
        m, X2p = 0, None
    else:
        X2 = as_host(X2)
        m, X2p = X2.shape[0], ptr(X2)
    out = np.empty(1 + ls.size)
    check(lib.gpb_kern_update_gradients_full_gower(_kind(kind), d, n, ptr(X), m, X2p, ptr(G), G.shape[1], float(variance), dptr(ls),
                                                   ls.size, flags.ctypes.data_as(_lib.c_int_p), dptr(rng), dptr(out), 0,
                                                   _lib.current_stream()), "update_gradients_full_gower")
    return out[0], out[1:].copy()


def kern_gradients_X(kind, dL_dK, X, X2, variance, lengthscale):
    """Stationary.gradients_X (stationary.py:271-278,354-364)."""
    lib = _lib.require_gpu()
    X = as_host(X)
    n, d = X.shape
    ls = _ls(lengthscale)
    G = as_host(dL_dK)
    if X2 is None:
        m, X2p = 0, None
        assert G.shape == (n, n)
    else:
        X2 = as_host(X2)
        m, X2p = X2.shape[0], ptr(X2)
        if G.shape == (1, m) and n != 1:  # broadcast row (core/gp.py:431-434 passes alpha^T)
            G = np.ascontiguousarray(np.broadcast_to(G, (n, m)))
        assert G.shape == (n, m)
    out = np.empty((n, d))
    check(lib.gpb_kern_gradients_X(_kind(kind), d, n, ptr(X), m, X2p, ptr(G), G.shape[1], float(variance), dptr(ls), ls.size, ptr(out),
                                   0, _lib.current_stream()), "gradients_X")
    return out


def pdinv(A, want=("Ai", "L", "Li", "logdet")):
    """pdinv without the jitter ladder (GPy/GPy/util/linalg.py:193-214); raises LinAlgError when not PD."""
    lib = _lib.require_gpu()
    A = as_host(A)
    n = A.shape[0]
    L = np.empty((n, n)) if "L" in want else None
    Ai = np.empty((n, n)) if "Ai" in want else None
    Li = np.empty((n, n)) if "Li" in want else None
    logdet = ctypes.c_double(0.0)
    rc = lib.gpb_pdinv(n, ptr(A), n, ptr(L), ptr(Ai), ptr(Li), ctypes.byref(logdet), 0, _lib.current_stream())
    return rc, Ai, L, Li, logdet.value


def potrs(L, B):
    """dpotrs(L, B, lower=1) (linalg.py:116-125)."""
    lib = _lib.require_gpu()
    L = as_host(L)
    Bc = as_host(B).copy()
    if Bc.ndim == 1:
        Bc = Bc[:, None]
    check(lib.gpb_potrs(L.shape[0], ptr(L), L.shape[0], ptr(Bc), Bc.shape[1], 0, _lib.current_stream()), "potrs")
    return Bc


def potri(L):
    """dpotri(L, lower=1) + symmetrify (linalg.py:127-145)."""
    lib = _lib.require_gpu()
    L = as_host(L)
    n = L.shape[0]
    Ai = np.empty((n, n))
    check(lib.gpb_potri(n, ptr(L), n, ptr(Ai), n, 0, _lib.current_stream()), "potri")
    return Ai


def dgemm(ta, tb, alpha, A, B, beta, C):
    """Device-only DMMA GEMM on CUDA torch tensors (row-major).  C is updated in place."""
    lib = _lib.require_gpu()
    assert is_torch(A) and is_torch(B) and is_torch(C)
    m, n = C.shape
    k = A.shape[0] if ta else A.shape[1]
    check(lib.gpb_dgemm(int(ta), int(tb), m, n, k, float(alpha), ptr(A), A.stride(0), ptr(B), B.stride(0), float(beta), ptr(C),
                        C.stride(0), _lib.current_stream()), "dgemm")
    return C


def ozaki_dgemm(ta, tb, alpha, A, B, beta, C, slices=0, tri_out=0, klo_mode=0, khi_mode=0, tri_a=0, tri_b=0):
    """Experimental: the same product through the int8 tensor cores (Ozaki scheme, csrc/gpb_ozaki.cu).  C is updated in place."""
    lib = _lib.require_gpu()
    assert is_torch(A) and is_torch(B) and is_torch(C)
    m, n = C.shape
    k = A.shape[0] if ta else A.shape[1]
    check(lib.gpb_ozaki_dgemm(int(ta), int(tb), m, n, k, float(alpha), ptr(A), A.stride(0), ptr(B), B.stride(0), float(beta), ptr(C),
                              C.stride(0), int(tri_out), int(klo_mode), int(khi_mode), int(tri_a), int(tri_b), int(slices),
                              _lib.current_stream()), "ozaki_dgemm")
    return C


def set_ozaki(min_n, slices=8):
    """Experimental: products of the factorisation with >= min_n rows go through the int8 engine (0 = off, the default).
    slices 1..8: balanced radix-256 digits per operand (S (S + 1) / 2 int8 products); 10..18: the modular mode, that many moduli
    (one int8 product each, CRT reconstruction; csrc/gpb_crt.cuh)."""
    check(_lib.load().gpb_set_ozaki(int(min_n), int(slices)), "set_ozaki")


def ozaki_fallback_count():
    """Fits of this process that the int8 engine's residual check sent back to the fp64 DMMA engine."""
    return int(_lib.load().gpb_ozaki_fallback_count())


def ozaki_crt_bits(nmod, k):
    """Bits per operand the modular mode of the int8 engine carries with nmod moduli at inner dimension k."""
    return int(_lib.load().gpb_ozaki_crt_bits(int(nmod), int(k)))


def gemm_config(cfg):
    """0 auto, 1 = 64x128, 2 = 64x64, 3 = 32x32 CTA tiles (tuning / tests)."""
    check(_lib.load().gpb_gemm_config(int(cfg)), "gemm_config")


def set_overlap(min_n):
    """Two-stream factorisation schedule: fork the T21 products of blocks with >= min_n rows (0 = single stream)."""
    check(_lib.load().gpb_set_overlap(int(min_n)), "set_overlap")


def profile_gemm(enable):
    check(_lib.load().gpb_profile_gemm(int(enable)), "profile_gemm")


def profile_gemm_collect():
    """-> (milliseconds in GEMM kernels, executed flops, launches) since the last collect."""
    ms, fl, n = ctypes.c_double(0), ctypes.c_double(0), ctypes.c_longlong(0)
    check(_lib.load().gpb_profile_gemm_collect(ctypes.byref(ms), ctypes.byref(fl), ctypes.byref(n)), "profile_gemm_collect")
    return ms.value, fl.value, n.value


def profile_gemm_last():
    """-> (milliseconds, executed flops) of the last recorded GEMM launch (call before profile_gemm_collect)."""
    ms, fl = ctypes.c_double(0), ctypes.c_double(0)
    check(_lib.load().gpb_profile_gemm_last(ctypes.byref(ms), ctypes.byref(fl)), "profile_gemm_last")
    return ms.value, fl.value


def launch_count():
    return int(_lib.load().gpb_launch_count())


class NativeModel(object):
    """Owner of one `gpb_model` handle: a GPRegression resident on the current CUDA device."""

    def __init__(self, kind, ard, d, p=1, n_cap=1024, cand_block=1024, use_torch_workspace=True):
        lib = _lib.require_gpu()
        self._lib = lib
        self.kind, self.ard, self.d, self.p = _kind(kind), bool(ard), int(d), int(p)
        self.n_cap, self.cand_block = int(n_cap), int(cand_block)
        self.nls = self.d if self.ard else 1
        self.n = 0
        self._ws = None
        self._theta = (1.0, np.ones(self.nls), 1.0, 0.0)
        # bumped by everything that changes what the resident posterior is (data, hyper-parameters, a fit, an append, an adopted
        # broadcast): PosteriorExact objects are views of this model and compare their own stamp with it on every access
        self.generation = 0
        ws_ptr, ws_bytes = None, 0
        if use_torch_workspace:
            try:
                import torch
                if torch.cuda.is_available():
                    ws_bytes = int(lib.gpb_model_workspace_bytes(self.n_cap, self.d, self.p, self.cand_block))
                    self._ws = torch.empty(ws_bytes + 256, dtype=torch.uint8, device="cuda")
                    base = self._ws.data_ptr()
                    ws_ptr = ctypes.c_void_p((base + 255) // 256 * 256)
            except ImportError:
                pass
        h = ctypes.c_void_p()
        check(lib.gpb_model_create(ctypes.byref(h), self.kind, int(self.ard), self.d, self.p, self.n_cap, self.cand_block, ws_ptr,
                                   ws_bytes, _lib.current_stream()), "model_create")
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._lib.gpb_model_destroy(self._h)
            self._h = None
            self._ws = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- data / parameters -------------------------------------------------------------------------------------------
    def set_data(self, X, Y):
        dev = is_torch(X)
        if not dev:
            X, Y = as_host(X), as_host(Y)
        n = X.shape[0]
        assert X.shape[1] == self.d and Y.shape == (n, self.p)
        self.generation += 1
        check(self._lib.gpb_model_set_data(self._h, n, ptr(X), ptr(Y), int(dev)), "set_data")
        self.n = n

    def set_gower(self, gower):
        """Matern52(Gower=True, space=...): gower = (continuous dims, discrete dims, ranges) or None to switch it off."""
        self.generation += 1
        if gower is None:
            check(self._lib.gpb_model_set_gower(self._h, 0, None, None), "set_gower")
            return
        flags, rng = _gower_args(gower, self.d)
        check(self._lib.gpb_model_set_gower(self._h, 1, flags.ctypes.data_as(_lib.c_int_p), dptr(rng)), "set_gower")

    def set_theta(self, variance, lengthscale, noise):
        ls = _ls(lengthscale)
        assert ls.size == self.nls
        check(self._lib.gpb_model_set_theta(self._h, float(variance), dptr(ls), float(noise)), "set_theta")
        old = self._theta
        if not (old[0] == float(variance) and old[2] == float(noise) and np.array_equal(old[1], ls)):   # bit-identical theta keeps
            self.generation += 1                                                                          # the factorisation (C side)
            self._theta = (float(variance), ls.copy(), float(noise), 0.0)

    def fit(self, want_grad=True, extra_jitter=0.0):
        """-> (info, log_marginal, grads or None); grads ordered [variance, lengthscale..., noise]."""
        out = np.zeros(3 + self.nls)
        self._theta = self._theta[:3] + (float(extra_jitter),)
        self.generation += 1
        rc = self._lib.gpb_model_fit(self._h, int(want_grad), float(extra_jitter), dptr(out))
        if rc < 0:
            check(rc, "fit")
        if rc > 0:
            return rc, None, None
        return 0, float(out[0]), (out[1:].copy() if want_grad else None)

    def engine_report(self):
        """-> (the current posterior came from the int8 engine, componentwise backward error of its solve or -1 if unchecked)."""
        used, res = ctypes.c_int(0), ctypes.c_double(0.0)
        check(self._lib.gpb_model_engine_report(self._h, ctypes.byref(used), ctypes.byref(res)), "engine_report")
        return bool(used.value), res.value

    def append(self, Xnew, Yall, want_grad=True):
        """Extend the fitted model by the rows Xnew (targets replaced by Yall) without refactorising: O(N^2 b).
        -> (info, log_marginal, grads or None) like fit()."""
        Xnew, Yall = as_host(Xnew), as_host(Yall)
        b = Xnew.shape[0]
        assert Xnew.shape[1] == self.d and Yall.shape == (self.n + b, self.p)
        out = np.zeros(3 + self.nls)
        self.generation += 1
        rc = self._lib.gpb_model_append(self._h, b, ptr(Xnew), ptr(Yall), 0, int(want_grad), dptr(out))
        if rc < 0:
            check(rc, "append")
        self.n += b
        if rc > 0:
            return rc, None, None
        return 0, float(out[0]), (out[1:].copy() if want_grad else None)

    def get(self, what, out=None):
        n = self.n
        shape = (n, self.p) if what == "alpha" else (n, n)
        dev = out is not None and is_torch(out)
        if out is None:
            out = np.empty(shape)
        ld = out.stride(0) if dev else out.shape[1]
        check(self._lib.gpb_model_get(self._h, what.encode(), ptr(out), int(ld), int(dev)), "get(%s)" % what)
        return out

    # -- prediction ----------------------------------------------------------------------------------------------------
    def predict(self, Xc, include_likelihood=True, want_var=True):
        Xc = as_host(Xc)
        mc = Xc.shape[0]
        mu = np.empty((mc, self.p))
        var = np.empty((mc, 1)) if want_var else None
        check(self._lib.gpb_model_predict(self._h, mc, ptr(Xc), int(include_likelihood), ptr(mu), ptr(var), 0), "predict")
        return mu, var

    def predict_full_cov(self, Xc, include_likelihood=True):
        Xc = as_host(Xc)
        mc = Xc.shape[0]
        mu = np.empty((mc, self.p))
        cov = np.empty((mc, mc))
        check(self._lib.gpb_model_predict_full_cov(self._h, mc, ptr(Xc), int(include_likelihood), ptr(mu), ptr(cov), 0),
              "predict_full_cov")
        return mu, cov

    def predictive_gradients(self, Xc, want_var=True):
        Xc = as_host(Xc)
        mc = Xc.shape[0]
        dmu = np.empty((mc, self.d, 1))
        dvar = np.empty((mc, self.d)) if want_var else None
        check(self._lib.gpb_model_predictive_gradients(self._h, mc, ptr(Xc), ptr(dmu), ptr(dvar), 0), "predictive_gradients")
        return dmu, dvar

    def fmin(self):
        v = ctypes.c_double(0.0)
        check(self._lib.gpb_model_fmin(self._h, ctypes.byref(v)), "fmin")
        return v.value

    def acquisition(self, acq, par, fmin, Xc, with_gradients=False, want_moments=False):
        """-> dict with f (mc,1) [, df (mc,d)] [, m, s (mc,1), dmdx, dsdx (mc,d)]; f = -acq like AcquisitionBase."""
        dev = is_torch(Xc)
        if dev:
            import torch
            mc = Xc.shape[0]
            new = lambda *s: torch.empty(*s, dtype=torch.float64, device=Xc.device)  # noqa: E731
        else:
            Xc = as_host(Xc)
            mc = Xc.shape[0]
            new = lambda *s: np.empty(s)  # noqa: E731
        r = {"f": new(mc, 1)}
        if with_gradients:
            r["df"] = new(mc, self.d)
        if want_moments:
            r["m"], r["s"] = new(mc, 1), new(mc, 1)
            if with_gradients:
                r["dmdx"], r["dsdx"] = new(mc, self.d), new(mc, self.d)
        aid = acq if isinstance(acq, int) else ACQ_IDS[acq]
        check(self._lib.gpb_model_acquisition(self._h, aid, float(par), float(fmin), mc, ptr(Xc), ptr(r["f"]), ptr(r.get("df")),
                                              ptr(r.get("m")), ptr(r.get("s")), ptr(r.get("dmdx")), ptr(r.get("dsdx")), int(dev)),
              "acquisition")
        return r

    def set_penalizers(self, transform, Xb, r, s):
        """AcquisitionLP.update_batches (LP.py:40-62); Xb None clears the penalisers (the log transform stays)."""
        t = {"none": 0, "softplus": 1}[transform] if isinstance(transform, str) else int(transform)
        if Xb is None:
            check(self._lib.gpb_model_set_penalizers(self._h, t, 0, None, None, None), "set_penalizers")
            return
        Xb, r, s = np.atleast_2d(as_host(Xb)), as_host(r).ravel(), as_host(s).ravel()   # run.py:1240-1253 passes one 1-D point
        assert Xb.shape == (r.size, self.d) and s.size == r.size
        check(self._lib.gpb_model_set_penalizers(self._h, t, r.size, ptr(Xb), ptr(r), ptr(s)), "set_penalizers")

    def acquisition_lp(self, acq, par, fmin, Xc, with_gradients=False):
        """AcquisitionLP.acquisition_function(_withGradients) (LP.py:70-140) -> f (mc,) [, df (mc, d)]."""
        Xc = as_host(Xc)
        mc = Xc.shape[0]
        f = np.empty(mc)
        df = np.empty((mc, self.d)) if with_gradients else None
        aid = acq if isinstance(acq, int) else ACQ_IDS[acq]
        check(self._lib.gpb_model_acquisition_lp(self._h, aid, float(par), float(fmin), mc, ptr(Xc), ptr(f), ptr(df), 0), "acquisition_lp")
        return (f, df) if with_gradients else f

    def acq_topk_full(self, acq, par, fmin, Xc, k, index_offset=0, with_gradients=True):
        """One pass: every candidate's f [and df] plus the k best -> (vals, idx, pts, f, df)."""
        dev = is_torch(Xc)
        if dev:
            import torch
            new = lambda *s: torch.empty(*s, dtype=torch.float64, device=Xc.device)  # noqa: E731
        else:
            Xc = as_host(Xc)
            new = lambda *s: np.empty(s)  # noqa: E731
        mc = Xc.shape[0]
        f = new(mc, 1)
        df = new(mc, self.d) if with_gradients else None
        vals, idx, pts = np.empty(k), np.empty(k, dtype=np.int64), np.empty((k, self.d))
        aid = acq if isinstance(acq, int) else ACQ_IDS[acq]
        check(self._lib.gpb_model_acq_topk_full(self._h, aid, float(par), float(fmin), mc, ptr(Xc), int(dev), int(k), int(index_offset),
                                                dptr(vals), idx.ctypes.data_as(_lib.c_ll_p), dptr(pts), ptr(f), ptr(df)), "acq_topk_full")
        return vals, idx, pts, f, df

    def acq_topk_dev(self, acq, par, fmin, Xc, k, index_offset=0, with_gradients=False, want_f=False, rows=None):
        """Device-resident variant without a host synchronisation: Xc a CUDA tensor -> rows (k, d + 2) CUDA tensor of
        [f, global index, coordinates] (empty slots [NaN, -1, NaN ...]) [, f (mc, 1), df (mc, d)], ready in stream order on the
        model's stream (issue the all-gather behind it: sharded.all_gather_topk_device)."""
        import torch
        assert is_torch(Xc)
        mc = Xc.shape[0]
        if rows is None:
            rows = torch.empty((k, self.d + 2), dtype=torch.float64, device=Xc.device)
        f = torch.empty((mc, 1), dtype=torch.float64, device=Xc.device) if (want_f or with_gradients) else None
        df = torch.empty((mc, self.d), dtype=torch.float64, device=Xc.device) if with_gradients else None
        aid = acq if isinstance(acq, int) else ACQ_IDS[acq]
        check(self._lib.gpb_model_acq_topk_dev(self._h, aid, float(par), float(fmin), mc, ptr(Xc), int(k), int(index_offset), ptr(rows),
                                               ptr(f), ptr(df)), "acq_topk_dev")
        return rows, f, df

    # -- multi-GPU state distribution ----------------------------------------------------------------------------------
    def state_tensor(self, what):
        """A float64 CUDA tensor VIEW of a resident array ("Li" = L^-1, "alpha", "L", "Wi") inside the model's torch workspace."""
        import torch
        if self._ws is None:
            raise _lib.GpbError("state_tensor needs the torch-owned workspace (use_torch_workspace=True)")
        p, cnt = ctypes.c_void_p(), ctypes.c_size_t()
        check(self._lib.gpb_model_state_ptr(self._h, what.encode(), ctypes.byref(p), ctypes.byref(cnt)), "state_ptr(%s)" % what)
        off = p.value - self._ws.data_ptr()
        assert 0 <= off and off + 8 * cnt.value <= self._ws.numel() and off % 8 == 0
        return self._ws[off:off + 8 * cnt.value].view(torch.float64)

    def theta(self):
        return self._theta

    def broadcast_state(self, src=0, group=None, parts=("Li", "alpha")):
        """the rank that fitted (src) broadcasts theta, alpha and L^-1 (optionally L and Ky^-1) over NCCL; the other
        ranks adopt them instead of refitting (GPModel.updateModel on every rank, gpmodel.py:78-93).  Every rank must hold the same
        data (set_data).  Collective: call on every rank of the group."""
        import torch
        import torch.distributed as dist
        rank = dist.get_rank(group)
        dev = self._ws.device
        head = torch.zeros(4 + self.nls, dtype=torch.float64, device=dev)
        if rank == src:
            v, ls, nz, jit = self._theta
            head[0], head[1], head[2], head[3] = float(self.n), v, nz, jit
            head[4:] = torch.from_numpy(np.asarray(ls, dtype=np.float64))
        dist.broadcast(head, src=src, group=group)
        h = head.cpu().numpy()
        assert int(h[0]) == self.n, "broadcast_state: rank %d holds %d points, the source %d" % (rank, self.n, int(h[0]))
        for what in parts:
            dist.broadcast(self.state_tensor(what), src=src, group=group)
        torch.cuda.current_stream().synchronize()
        if rank != src:
            ls = np.ascontiguousarray(h[4:])
            mask = (1 if "L" in parts else 0) | (2 if "Wi" in parts else 0)
            self.generation += 1
            check(self._lib.gpb_model_adopt_state(self._h, float(h[1]), dptr(ls), float(h[2]), float(h[3]), mask), "adopt_state")
            self._theta = (float(h[1]), ls.copy(), float(h[2]), float(h[3]))

    def acq_topk(self, acq, par, fmin, Xc, k, index_offset=0):
        dev = is_torch(Xc)
        if not dev:
            Xc = as_host(Xc)
        mc = Xc.shape[0]
        vals = np.empty(k)
        idx = np.empty(k, dtype=np.int64)
        pts = np.empty((k, self.d))
        aid = acq if isinstance(acq, int) else ACQ_IDS[acq]
        check(self._lib.gpb_model_acq_topk(self._h, aid, float(par), float(fmin), mc, ptr(Xc), int(dev), int(k), int(index_offset),
                                           dptr(vals), idx.ctypes.data_as(_lib.c_ll_p), dptr(pts)), "acq_topk")
        return vals, idx, pts
