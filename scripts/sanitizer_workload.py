"""Unit-size pass over EVERY kernel family of libgpb200.so, for compute-sanitizer (scripts/run_sanitizer.sh):
stateless Kern / linalg entry points, the resident-model path (one-kernel tiny fit, general path, CUDA-graph replay, two-stream
factorisation, append), every predictive route (skinny M <= 8, blocked, full covariance, gradients), EI / LCB / LP epilogues,
top-k (host and device variants), the Gower kernel, and the int8 tensor-core engine (tcgen05 + TMEM + TMA multicast over CTA
pairs) in digit and modular mode, alone and inside a factorisation.  Results are checked loosely so that a wrong answer fails too.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gaussian_process_optimization_b200 import native  # noqa: E402

SMALL = os.environ.get("GPB_SANITIZER_SMALL", "0") == "1"      # racecheck: ~100x slower, keep the sizes minimal


def data(n, d, seed=0):
    rs = np.random.RandomState(seed)
    X = rs.uniform(0, 1, (n, d))
    Y = np.sin(X.sum(1))[:, None] + 0.05 * rs.randn(n, 1)
    return X, (Y - Y.mean()) / Y.std()


def main():
    import torch
    rs = np.random.RandomState(1)
    # ---- stateless Kern / linalg ----
    X, Y = data(150, 4)
    ls = np.array([0.5, 0.6, 0.7, 0.8])
    for kind in ("rbf", "mat52"):
        K = native.kern_K(kind, X, None, 1.2, ls)
        assert np.allclose(np.diag(K), 1.2)
        native.kern_K(kind, X[:40], X, 1.2, ls)
        native.kern_update_gradients_full(kind, rs.randn(150, 150), X, None, 1.2, ls)
        native.kern_update_gradients_full(kind, rs.randn(40, 150), X[:40], X, 1.2, ls)
        native.kern_gradients_X(kind, rs.randn(40, 150), X[:40], X, 1.2, ls)
        native.kern_gradients_X(kind, rs.randn(150, 150), X, None, 1.2, ls)
    n = 300 if SMALL else 600
    B = rs.randn(n, n + 3)
    A = B @ B.T + 0.5 * n * np.eye(n)
    rc, Ai, L, Li, logdet = native.pdinv(A)
    assert rc == 0 and np.allclose(Ai @ A, np.eye(n), atol=1e-8)
    native.potrs(L, rs.randn(n, 2))
    native.potri(L)
    # ---- resident model: tiny (N <= 128), general, graph replay ----
    for n, d in ((100, 3), (300, 4)) + (() if SMALL else ((1100, 6),)):
        X, Y = data(n, d, seed=n)
        ls = 0.5 + 0.1 * np.arange(d)
        for kind in ("rbf", "mat52"):
            m = native.NativeModel(kind, True, d, 1, n_cap=n + 200, cand_block=256)
            m.set_data(X, Y)
            for rep in range(3):                              # the third evaluation of a shape replays the captured graph
                m.set_theta(1.0 + 0.01 * rep, ls, 1e-2)
                info, logL, g = m.fit(True)
                assert info == 0 and np.isfinite(logL) and np.all(np.isfinite(g))
            fmin = m.fmin()
            for mc in (1, 3, 8, 200):
                Xc = rs.uniform(0, 1, (mc, d))
                m.predict(Xc)
                m.predictive_gradients(Xc)
                for acq, par in (("EI", 0.01), ("LCB", 2.0)):
                    r = m.acquisition(acq, par, fmin, Xc, with_gradients=True, want_moments=True)
                    assert np.all(np.isfinite(r["f"])) and np.all(np.isfinite(r["df"]))
            m.predict_full_cov(rs.uniform(0, 1, (50, d)))
            Xc = rs.uniform(0, 1, (300, d))
            m.acq_topk("EI", 0.01, fmin, Xc, 5)
            m.acq_topk_full("LCB", 2.0, 0.0, Xc, 5)
            rows, f, df = m.acq_topk_dev("EI", 0.01, fmin, torch.from_numpy(Xc).cuda(), 5, with_gradients=True)
            torch.cuda.synchronize()
            m.set_penalizers("none", Xc[:3], np.array([0.3, 0.2, 0.1]), np.array([0.05, 0.04, 0.03]))
            m.acquisition_lp("EI", 0.01, fmin, Xc[:20], with_gradients=True)
            m.set_penalizers("softplus", Xc[:2], np.array([0.3, 0.2]), np.array([0.05, 0.04]))
            m.acquisition_lp("LCB", 2.0, 0.0, Xc[:20], with_gradients=True)
            m.set_penalizers("none", None, None, None)
            # append 1 and 60 rows
            for b in (1, 60):
                Xn = rs.uniform(0, 1, (b, d))
                Xall = np.vstack([X, Xn])
                Yall = np.vstack([Y, rs.randn(b, 1) * 0.1])
                info, logL, g = m.append(Xn, Yall)
                assert info == 0 and np.isfinite(logL)
                X, Y = Xall, Yall
            m.get("Wi")
            m.get("L")
            m.get("dL_dK")
            X, Y = data(n, d, seed=n)
            m.close()
    # ---- Gower kernel ----
    X, Y = data(150, 4, seed=9)
    X[:, 3] = np.round(X[:, 3] * 3)
    gower = ([0, 1, 2], [3], [1.0, 1.0, 1.0])
    native.kern_K_gower("mat52", X, None, 1.1, gower)
    native.kern_update_gradients_full_gower("mat52", rs.randn(150, 150), X, None, 1.1, np.full(4, 0.7), gower)
    m = native.NativeModel("mat52", True, 4, 1, n_cap=256, cand_block=128)
    m.set_data(X, Y)
    m.set_gower(gower)
    m.set_theta(1.1, np.full(4, 0.7), 1e-2)
    assert m.fit(True)[0] == 0
    m.acquisition("EI", 0.01, m.fmin(), X[:10] + 0.01, with_gradients=True)
    m.close()
    # ---- DMMA GEMM engine, every tile configuration ----
    for cfg in (1, 2, 3, 9):
        native.gemm_config(cfg)
        A = torch.randn(256, 272, dtype=torch.float64, device="cuda")
        Bm = torch.randn(384, 272, dtype=torch.float64, device="cuda")
        C = torch.zeros(256, 384, dtype=torch.float64, device="cuda")
        native.dgemm(0, 0, 1.0, A, Bm, 0.0, C)
        torch.cuda.synchronize()
        assert torch.allclose(C, A @ Bm.T, atol=1e-10)
    native.gemm_config(0)
    # ---- int8 tensor-core engine: tcgen05 / TMEM / TMA multicast (CTA pairs need two 256-row tiles with a common k-range) ----
    mm = 512
    A = torch.randn(mm, 384, dtype=torch.float64, device="cuda")
    Bm = torch.randn(mm, 384, dtype=torch.float64, device="cuda")
    for slices in (8, 16):
        C = torch.zeros(mm, mm, dtype=torch.float64, device="cuda")
        native.ozaki_dgemm(0, 0, 1.0, A, Bm, 0.0, C, slices=slices)
        torch.cuda.synchronize()
        assert torch.allclose(C, A @ Bm.T, atol=1e-10)
        C = torch.zeros(mm, mm, dtype=torch.float64, device="cuda")
        native.ozaki_dgemm(1, 1, 1.0, A.T.contiguous(), Bm.T.contiguous(), 0.0, C, slices=slices)
        torch.cuda.synchronize()
        assert torch.allclose(C, A @ Bm.T, atol=1e-10)
    if not SMALL:
        X, Y = data(1100, 5, seed=4)
        for slices in (8, 16):
            native.set_ozaki(256, slices)
            m = native.NativeModel("rbf", True, 5, 1, n_cap=1152, cand_block=1024)
            m.set_data(X, Y)
            m.set_theta(1.0, np.full(5, 0.6), 1e-2)
            info, logL, g = m.fit(True)
            assert info == 0 and np.isfinite(logL)
            m.acq_topk_full("EI", 0.01, m.fmin(), rs.uniform(0, 1, (1024, 5)), 5)
            m.close()
        native.set_ozaki(0)
    print("sanitizer workload OK; kernels launched:", native.launch_count())


if __name__ == "__main__":
    main()
