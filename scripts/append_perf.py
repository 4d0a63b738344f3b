"""Device timings of the incremental model update (gpb_model_append) against the full refit it replaces.
Not the bench contract: wall clock around one C-ABI call with host inputs, best of a few repeats."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from bench import synth  # noqa: E402
from gaussian_process_optimization_b200 import native  # noqa: E402

out = {}
cases = [("rbf", 16384 - 256, 16, (1, 64, 64, 127, 1)), ("mat52", 32768 - 192, 20, (64, 64, 64))]
if len(sys.argv) > 1:
    cases = [("rbf", int(sys.argv[1]), 16, (1, 64, 64))]
for kind, n0, D, steps in cases:
    ntot = n0 + sum(steps)
    X, Y, ls = synth(ntot, D)
    m = native.NativeModel(kind, True, D, 1, n_cap=ntot, cand_block=1024)
    m.set_data(X[:n0], Y[:n0])
    m.set_theta(1.0, ls, 1e-2)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    info, logL, g = m.fit(True)
    torch.cuda.synchronize()
    t_full = time.perf_counter() - t0
    t0 = time.perf_counter()
    info, logL, g = m.fit(True)
    torch.cuda.synchronize()
    t_full = min(t_full, time.perf_counter() - t0)
    print("%s N=%d D=%d full fit (value+gradient): %.2f ms" % (kind, n0, D, t_full * 1e3), flush=True)
    rows = []
    n = n0
    for i, b in enumerate(steps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        info, logL, g = m.append(X[n:n + b], Y[:n + b], want_grad=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        n += b
        assert info == 0
        rows.append({"n_after": n, "b": b, "ms": dt * 1e3, "logL": logL})
        print("  append b=%d -> N=%d (value+gradient): %.2f ms" % (b, n, dt * 1e3), flush=True)
    # the same end state from scratch
    ref = native.NativeModel(kind, True, D, 1, n_cap=ntot, cand_block=1024) if ntot <= 20000 else None
    if ref is not None:
        ref.set_data(X[:n], Y[:n])
        ref.set_theta(1.0, ls, 1e-2)
        info, l_ref, g_ref = ref.fit(True)
        print("  end state vs full refit: |dlogL|/|logL| = %.2e, max rel grad diff = %.2e" %
              (abs(logL - l_ref) / abs(l_ref), np.max(np.abs(g - g_ref) / np.abs(g_ref))), flush=True)
        rows.append({"rel_logL": abs(logL - l_ref) / abs(l_ref), "rel_grad": float(np.max(np.abs(g - g_ref) / np.abs(g_ref)))})
        ref.close()
    out["%s_N%d_D%d" % (kind, n0, D)] = {"full_fit_ms": t_full * 1e3, "appends": rows}
    m.close()
json.dump(out, open("gpurun_out/append_perf.json", "w"), indent=1)
