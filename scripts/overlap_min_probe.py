import sys, json, numpy as np, torch
sys.path.insert(0, ".")
from bench import synth
from gaussian_process_optimization_b200 import native
for N, D in ((1024, 8), (4096, 8), (16384, 16)):
    X, Y, ls = synth(N, D)
    m = native.NativeModel("rbf", True, D, 1, n_cap=N, cand_block=128)
    m.set_data(torch.from_numpy(X).cuda(), torch.from_numpy(Y).cuda())
    for ov in (512, 256, 128):
        native.set_overlap(ov)
        ts = []
        for i in range(14):
            m.set_theta(1.0 + 1e-3 * (i % 5), ls, 1e-2)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); info, logL, g = m.fit(True); b.record(); torch.cuda.synchronize()
            if i >= 4: ts.append(a.elapsed_time(b))
        print(N, "overlap_min", ov, "ms %.4f" % min(ts), flush=True)
    native.set_overlap(512)
    m.close()
