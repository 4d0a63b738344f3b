"""Red-zone tests: the pool's compute-sanitizer is closed (profiles/r2_sanitizer.md), so out-of-bounds accesses are hunted with
guards of our own.  Every device-pointer entry point is given operands that sit INSIDE larger allocations whose surroundings are
poisoned: a write outside the declared extent destroys the sentinel pattern, a read outside it drags a NaN into the result.  Sizes
are deliberately ragged (not multiples of the 64 / 128 / 256 tile edges) and strides larger than the row length.  Races show up
as run-to-run differences: every call is repeated and must reproduce bit for bit.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

native = pytest.importorskip("gaussian_process_optimization_b200.native")
from gaussian_process_optimization_b200 import _lib  # noqa: E402

SENT = -7.25e301


def _guarded(torch, rows, cols, pad_r=3, pad_c=5, fill=None):
    """(whole, view): `view` is the rows x cols interior of a sentinel-filled matrix (row stride cols + 2 pad_c)."""
    whole = torch.full((rows + 2 * pad_r, cols + 2 * pad_c), SENT, dtype=torch.float64, device="cuda")
    view = whole[pad_r:pad_r + rows, pad_c:pad_c + cols]
    if fill is not None:
        view.copy_(fill)
    return whole, view


def _guards_intact(torch, whole, rows, cols, pad_r=3, pad_c=5):
    mask = torch.ones_like(whole, dtype=torch.bool)
    mask[pad_r:pad_r + rows, pad_c:pad_c + cols] = False
    return bool((whole[mask] == SENT).all())


def _poisoned_input(torch, a, pad_r=2, pad_c=4):
    """A copy of the host matrix `a` inside a NaN-filled device allocation; contiguous rows are required by the C ABI for X-like
    inputs, so only rows before / after are poisoned there (pad_c = 0)."""
    rows, cols = a.shape
    whole = torch.full((rows + 2 * pad_r, cols + 2 * pad_c), float("nan"), dtype=torch.float64, device="cuda")
    view = whole[pad_r:pad_r + rows, pad_c:pad_c + cols]
    view.copy_(torch.from_numpy(a))
    return whole, view


@pytest.mark.parametrize("n,m,d", [(1, 1, 1), (63, 65, 3), (129, 257, 16), (300, 77, 33)])
@pytest.mark.parametrize("kind", ["rbf", "mat52"])
def test_kern_K_device_pointers_stay_inside_their_extents(n, m, d, kind):
    import torch
    lib = _lib.require_gpu()
    rs = np.random.RandomState(n + m + d)
    X, X2 = rs.uniform(0, 1, (n, d)), rs.uniform(0, 1, (m, d))
    ls = np.ascontiguousarray(0.5 + 0.1 * np.arange(d))
    wx, vx = _poisoned_input(torch, X, pad_c=0)
    wx2, vx2 = _poisoned_input(torch, X2, pad_c=0)
    ref_sq, ref_rect = native.kern_K(kind, X, None, 1.3, ls), native.kern_K(kind, X, X2, 1.3, ls)
    for rep in range(2):
        wk, vk = _guarded(torch, n, n)
        _lib.check(lib.gpb_kern_K(_lib.KIND_IDS[kind], d, n, _lib.c_void_p(vx.data_ptr()), 0, None, 1.3, _lib.dptr(ls), d,
                                  _lib.c_void_p(vk.data_ptr()), wk.stride(0), 1, _lib.current_stream()), "kern_K")
        assert _guards_intact(torch, wk, n, n) and np.array_equal(vk.cpu().numpy(), ref_sq)
        wk, vk = _guarded(torch, n, m)
        _lib.check(lib.gpb_kern_K(_lib.KIND_IDS[kind], d, n, _lib.c_void_p(vx.data_ptr()), m, _lib.c_void_p(vx2.data_ptr()), 1.3,
                                  _lib.dptr(ls), d, _lib.c_void_p(vk.data_ptr()), wk.stride(0), 1, _lib.current_stream()), "kern_K")
        assert _guards_intact(torch, wk, n, m) and np.array_equal(vk.cpu().numpy(), ref_rect)


@pytest.mark.parametrize("n", [1, 127, 129, 300, 641])
def test_pdinv_device_pointers_stay_inside_their_extents(n):
    import torch
    lib = _lib.require_gpu()
    rs = np.random.RandomState(n)
    B = rs.randn(n, n + 2)
    A = B @ B.T + 0.5 * n * np.eye(n)
    wa, va = _poisoned_input(torch, A)
    outs = []
    for rep in range(2):
        # the C ABI takes dense n x n outputs: hand it the contiguous interiors of row-padded allocations
        Ld = torch.full((n + 6, n), SENT, dtype=torch.float64, device="cuda")
        Ai = torch.full((n + 6, n), SENT, dtype=torch.float64, device="cuda")
        Li = torch.full((n + 6, n), SENT, dtype=torch.float64, device="cuda")
        logdet = _lib.ctypes.c_double(0.0)
        rc = lib.gpb_pdinv(n, _lib.c_void_p(va.data_ptr()), wa.stride(0), _lib.c_void_p(Ld[3:].data_ptr()), _lib.c_void_p(Ai[3:].data_ptr()),
                           _lib.c_void_p(Li[3:].data_ptr()), _lib.ctypes.byref(logdet), 1, _lib.current_stream())
        assert rc == 0
        for t in (Ld, Ai, Li):
            assert bool((t[:3] == SENT).all()) and bool((t[3 + n:] == SENT).all())
            assert bool(torch.isfinite(t[3:3 + n]).all())
        outs.append((Ld[3:3 + n].cpu().numpy(), Ai[3:3 + n].cpu().numpy(), Li[3:3 + n].cpu().numpy(), logdet.value))
    Lr = np.linalg.cholesky(A)
    np.testing.assert_allclose(outs[0][0], Lr, rtol=1e-9, atol=1e-12 * np.abs(Lr).max())
    for a, b in zip(outs[0], outs[1]):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("n,d,mc", [(1, 1, 1), (100, 3, 7), (129, 5, 130), (700, 16, 300)])
def test_model_device_outputs_stay_inside_their_extents(n, d, mc):
    import torch
    rs = np.random.RandomState(n + mc)
    X = rs.uniform(0, 1, (n, d))
    Y = np.sin(X.sum(1))[:, None] + 0.05 * rs.randn(n, 1)
    wx, vx = _poisoned_input(torch, X, pad_c=0)
    wy, vy = _poisoned_input(torch, Y, pad_c=0)
    nm = native.NativeModel("mat52", True, d, 1, n_cap=n, cand_block=128)
    nm.set_data(vx, vy)
    nm.set_theta(1.2, 0.5 + 0.05 * np.arange(d), 1e-2)
    info, logL, g = nm.fit(True)
    assert info == 0 and np.isfinite(logL) and np.all(np.isfinite(g))
    fmin = nm.fmin()
    Xc = rs.uniform(0, 1, (mc, d))
    wc, vc = _poisoned_input(torch, Xc, pad_c=0)
    ref = nm.acquisition("EI", 0.01, fmin, Xc, with_gradients=True, want_moments=True)
    k = min(5, mc)
    for rep in range(2):
        bufs = {key: torch.full((mc + 4, w), SENT, dtype=torch.float64, device="cuda")
                for key, w in (("f", 1), ("df", d), ("m", 1), ("s", 1), ("dmdx", d), ("dsdx", d))}
        p = {key: _lib.c_void_p(b[2:].data_ptr()) for key, b in bufs.items()}
        _lib.check(nm._lib.gpb_model_acquisition(nm._h, 0, 0.01, fmin, mc, _lib.c_void_p(vc.data_ptr()), p["f"], p["df"], p["m"], p["s"],
                                                 p["dmdx"], p["dsdx"], 1), "acquisition")
        for key, b in bufs.items():
            assert bool((b[:2] == SENT).all()) and bool((b[2 + mc:] == SENT).all()), key
            assert np.array_equal(b[2:2 + mc].cpu().numpy(), ref[key]), key
        rows = torch.full((k + 2, d + 2), SENT, dtype=torch.float64, device="cuda")
        fd = torch.full((mc + 2, 1), SENT, dtype=torch.float64, device="cuda")
        _lib.check(nm._lib.gpb_model_acq_topk_dev(nm._h, 0, 0.01, fmin, mc, _lib.c_void_p(vc.data_ptr()), k, 0, _lib.c_void_p(rows[1:].data_ptr()),
                                                  _lib.c_void_p(fd[1:].data_ptr()), None), "acq_topk_dev")
        torch.cuda.synchronize()
        assert bool((rows[0] == SENT).all()) and bool((rows[k + 1] == SENT).all()) and bool((fd[0] == SENT).all()) and bool((fd[-1] == SENT).all())
        order = np.argsort(ref["f"].ravel(), kind="stable")[:k]
        assert np.array_equal(rows[1:1 + k, 1].cpu().numpy(), order.astype(float))
        assert np.array_equal(fd[1:1 + mc].cpu().numpy(), ref["f"])
    # accessors into guarded, strided destinations
    for what in ("L", "Li", "Wi", "K", "dL_dK"):
        w, v = _guarded(torch, n, n)
        _lib.check(nm._lib.gpb_model_get(nm._h, what.encode(), _lib.c_void_p(v.data_ptr()), w.stride(0), 1), "get")
        assert _guards_intact(torch, w, n, n) and bool(torch.isfinite(v).all()), what
    nm.close()


@pytest.mark.parametrize("slices", [8, 16])
@pytest.mark.parametrize("m,n,k", [(128, 128, 128), (384, 256, 640)])
def test_engines_write_only_their_output_tiles(slices, m, n, k):
    """DMMA engine and int8 engine on operands / outputs embedded in larger, poisoned allocations (leading dimensions > extents)."""
    import torch
    g = torch.Generator(device="cpu").manual_seed(m + n + k)
    A, B = torch.randn(m, k, generator=g, dtype=torch.float64), torch.randn(n, k, generator=g, dtype=torch.float64)
    wa, va = _poisoned_input(torch, A.numpy(), pad_r=1, pad_c=16)
    wb, vb = _poisoned_input(torch, B.numpy(), pad_r=1, pad_c=16)
    lib = _lib.require_gpu()
    ref = (A @ B.T).numpy()
    for engine in ("dmma", "int8"):
        prev = None
        for rep in range(2):
            wc, vc = _guarded(torch, m, n, pad_r=2, pad_c=16)
            if engine == "dmma":
                rc = lib.gpb_dgemm(0, 0, m, n, k, 1.0, _lib.c_void_p(va.data_ptr()), wa.stride(0), _lib.c_void_p(vb.data_ptr()), wb.stride(0),
                                   0.0, _lib.c_void_p(vc.data_ptr()), wc.stride(0), _lib.current_stream())
            else:
                rc = lib.gpb_ozaki_dgemm(0, 0, m, n, k, 1.0, _lib.c_void_p(va.data_ptr()), wa.stride(0), _lib.c_void_p(vb.data_ptr()),
                                         wb.stride(0), 0.0, _lib.c_void_p(vc.data_ptr()), wc.stride(0), 0, 0, 0, 0, 0, slices,
                                         _lib.current_stream())
            _lib.check(rc, engine)
            torch.cuda.synchronize()
            assert _guards_intact(torch, wc, m, n, pad_r=2, pad_c=16), engine
            out = vc.cpu().numpy()
            np.testing.assert_allclose(out, ref, rtol=0, atol=1e-11 * np.abs(ref).max())
            assert prev is None or np.array_equal(prev, out)
            prev = out


def test_misuse_returns_error_codes_and_leaves_the_model_usable():
    """Wrong call order, NULL pointers, sizes beyond the capacity, unknown names: every entry point answers with a negative code and
    a message (gpb_last_error), never with a crash, and the model gives the same answers afterwards as before.  Non-finite inputs
    travel like in the reference: a NaN in X makes the factorisation fail with LAPACK's info > 0 (jitchol's LinAlgError,
    util/linalg.py:56-75), a NaN candidate gives a NaN score and never wins a top-k slot."""
    import ctypes

    from gaussian_process_optimization_b200 import _lib
    from gaussian_process_optimization_b200._lib import ptr, dptr
    lib = _lib.load()
    rs = np.random.RandomState(77)
    N, D = 60, 3
    X, Y = rs.uniform(0, 1, (N, D)), rs.randn(N, 1)
    ls = np.array([0.5, 0.6, 0.7])
    m = native.NativeModel("rbf", True, D, 1, n_cap=64, cand_block=128)
    out = np.zeros(3 + D)
    Xc = rs.uniform(0, 1, (5, D))
    mu, var = np.empty((5, 1)), np.empty((5, 1))

    def bad(rc, needle=None):
        assert rc < 0, rc
        msg = lib.gpb_last_error().decode()
        assert msg and (needle is None or needle in msg), msg

    bad(lib.gpb_model_fit(m._h, 1, 0.0, dptr(out)), "set_data")                                  # fit before data
    bad(lib.gpb_model_predict(m._h, 5, ptr(Xc), 1, ptr(mu), ptr(var), 0), "fitted")              # predict before fit
    bad(lib.gpb_model_set_data(m._h, 65, ptr(rs.randn(65, D)), ptr(rs.randn(65, 1)), 0))          # beyond n_cap
    bad(lib.gpb_model_set_data(m._h, N, None, ptr(Y), 0), "NULL")
    m.set_data(X, Y)
    m.set_theta(1.0, ls, 1e-2)
    # hyper-parameters outside their domain: like the reference, where they reach LAPACK and come back as "not positive definite"
    # (a LinAlgError, which paramz's optimiser loop survives), the fit answers GPB_ERR_DOMAIN -> numpy.linalg.LinAlgError
    for v_bad, ls_bad in ((-1.0, ls), (1.0, np.array([0.5, 0.0, 0.7])), (float("nan"), ls), (1.0, np.array([0.5, np.nan, 0.7]))):
        m.set_theta(v_bad, ls_bad, 1e-2)
        assert lib.gpb_model_fit(m._h, 1, 0.0, dptr(out)) == _lib.ERR_DOMAIN, (v_bad, ls_bad)
        with pytest.raises(np.linalg.LinAlgError):
            m.fit(True)
    m.set_theta(1.0, ls, 1e-2)
    bad(lib.gpb_model_fit(m._h, 1, 0.0, None), "NULL")
    info, logL, g = m.fit(True)
    assert info == 0
    ref = m.acquisition("EI", 0.01, m.fmin(), Xc, with_gradients=True)
    bad(lib.gpb_model_predict(m._h, -1, ptr(Xc), 1, ptr(mu), ptr(var), 0), "negative")
    bad(lib.gpb_model_predict(m._h, 5, None, 1, ptr(mu), ptr(var), 0), "NULL")
    big = rs.uniform(0, 1, (129, D))
    cov = np.empty((129, 129))
    bad(lib.gpb_model_predict_full_cov(m._h, 129, ptr(big), 1, ptr(np.empty((129, 1))), ptr(cov), 0), "candidate block")
    bad(lib.gpb_model_get(m._h, b"no_such_array", ptr(np.empty((N, N))), N, 0))
    bad(lib.gpb_model_append(m._h, 5, ptr(rs.randn(5, D)), ptr(rs.randn(N + 5, 1)), 0, 1, dptr(out)))   # 60 + 5 > n_cap = 64
    bad(lib.gpb_model_acquisition(m._h, 7, 0.01, 0.0, 5, ptr(Xc), ptr(mu), None, None, None, None, None, 0))   # unknown acquisition
    bad(lib.gpb_kern_K(9, D, 5, ptr(Xc), 0, None, 1.0, dptr(ls), D, ptr(np.empty((5, 5))), 5, 0, None))        # unknown kernel
    bad(lib.gpb_kern_K(0, 97, 5, ptr(rs.randn(5, 97)), 0, None, 1.0, dptr(np.ones(97)), 97, ptr(np.empty((5, 5))), 5, 0, None))   # d <= 96
    bad(lib.gpb_kern_gradients_X(0, 65, 5, ptr(rs.randn(5, 65)), 0, None, ptr(rs.randn(5, 5)), 5, 1.0, dptr(np.ones(65)), 65,
                                 ptr(np.empty((5, 65))), 0, None), "64")                                                        # d <= 64
    bad(lib.gpb_pdinv(0, ptr(np.eye(1)), 1, None, None, None, ctypes.byref(ctypes.c_double()), 0, None))
    h = ctypes.c_void_p()
    bad(lib.gpb_model_create(ctypes.byref(h), 0, 1, 0, 1, 64, 128, None, 0, None))                # zero input dimension
    bad(lib.gpb_model_create(ctypes.byref(h), 0, 1, 3, 1, 0, 128, None, 0, None))                 # zero capacity
    # nothing above disturbed the model
    again = m.acquisition("EI", 0.01, m.fmin(), Xc, with_gradients=True)
    assert np.array_equal(ref["f"], again["f"]) and np.array_equal(ref["df"], again["df"])
    info2, logL2, g2 = m.fit(True)
    assert info2 == 0 and logL2 == logL and np.array_equal(g, g2)
    # non-finite values
    Xn = Xc.copy()
    Xn[2, 1] = np.nan
    r = m.acquisition("EI", 0.01, m.fmin(), Xn, with_gradients=False)
    assert np.isnan(r["f"][2, 0]) and np.all(np.isfinite(np.delete(r["f"], 2, axis=0)))
    vals, idx, pts = m.acq_topk("EI", 0.01, m.fmin(), Xn, 5)
    assert 2 not in list(idx) and list(idx).count(-1) == 1
    Xbad = X.copy()
    Xbad[7, 0] = np.nan
    m.set_data(Xbad, Y)
    m.set_theta(1.0, ls, 1e-2)
    info, _, _ = m.fit(True)
    assert info > 0                                    # "not positive definite": the reference's jitchol raises here as well
    m.set_data(X, Y)
    m.set_theta(1.0, ls, 1e-2)
    info3, logL3, g3 = m.fit(True)
    assert info3 == 0 and logL3 == logL and np.array_equal(g, g3)
    m.close()
