"""How far is the CUDA path from the reference-generated full-size vectors (tests/golden/fullsize/*.npz)?  Prints / writes the
measured differences per quantity, per case and per engine (fp64 DMMA; int8 engine forced on every product >= 256 rows with 16 and
18 moduli and with 7 and 8 digits).  tests/test_gpu_fullsize_golden.py holds the bars; this is the evidence behind them.

    python scripts/fullsize_parity_report.py [out.json]
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from gaussian_process_optimization_b200 import native  # noqa: E402
import test_gpu_fullsize_golden as T  # noqa: E402


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def relwise(a, b):
    a, b = np.asarray(a, float).ravel(), np.asarray(b, float).ravel()
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


def run(name, engine):
    g = T._load(name)
    X, Y = T._synth(g["N"], g["D"], g["data_seed"])
    native.set_ozaki(*engine[1]) if engine[1] else native.set_ozaki(0)
    out = {"cond_bound": g["cond_bound"], "cond_factor": T._cond_factor(g)}
    try:
        m = native.NativeModel(g["kind"], True, g["D"], 1, n_cap=g["N"], cand_block=2048)
        m.set_data(X, Y)
        m.set_theta(g["variance"], g["lengthscale"], g["noise"])
        info, logL, grads = m.fit(True)
        out["info"] = int(info)
        if info != 0:
            return out
        out["logL_rel"] = abs(logL - g["logL"]) / abs(g["logL"])
        out["grads_rel_to_max"] = rel(grads, g["grads"])
        out["grads_relwise"] = relwise(grads, g["grads"])
        out["alpha_rel_to_max"] = rel(m.get("alpha")[g["rows"], 0], g["alpha_rows"])
        Xc = np.random.RandomState(g["cand_seed"]).uniform(0, 1, (2 ** 16, g["D"]))
        fmin = m.fmin()
        out["fmin_rel"] = abs(fmin - g["fmin"]) / abs(g["fmin"])
        vals, idx, pts, f, _ = m.acq_topk_full("EI", 0.01, fmin, Xc, 5, with_gradients=False)
        out["ei_top5_identical"] = bool(np.array_equal(idx, g["ei_top5_idx"]))
        out["ei_top5_val_relwise"] = relwise(vals, g["ei_top5_val"])
        out["ei_f_rel_to_max"] = rel(f.ravel(), g["ei_f"])
        big = np.abs(g["ei_f"]) > 1e-6 * np.abs(g["ei_f"]).max()
        out["ei_f_relwise_where_above_1e-6_of_max"] = relwise(f.ravel()[big], g["ei_f"][big])
        Xg = Xc[:g["ei_g_f"].size]
        r = m.acquisition("EI", 0.01, fmin, Xg, with_gradients=True, want_moments=True)
        out["mean_rel_to_max"] = rel(r["m"].ravel(), g["gpm_m"])
        out["sd_relwise"] = relwise(r["s"].ravel(), g["gpm_s"])
        out["dmdx_rel_to_max"] = rel(r["dmdx"], g["gpm_dmdx"])
        out["dsdx_rel_to_max"] = rel(r["dsdx"], g["gpm_dsdx"])
        out["ei_df_rel_to_max"] = rel(r["df"], g["ei_g_df"])
        Xl = Xc[:g["lcb_f"].size]
        v2, i2, _, f2, _ = m.acq_topk_full("LCB", 2.0, 0.0, Xl, 5, with_gradients=False)
        out["lcb_top5_identical"] = bool(np.array_equal(i2, g["lcb_top5_idx"]))
        out["lcb_f_rel_to_max"] = rel(f2.ravel(), g["lcb_f"])
        r2 = m.acquisition("LCB", 2.0, 0.0, Xg, with_gradients=True)
        out["lcb_df_rel_to_max"] = rel(r2["df"], g["lcb_g_df"])
        out["oracle_vs_ref"] = {k: float(g[k]) for k in g if k.startswith("oracle_vs_ref")}
        m.close()
    except Exception as exc:
        out["error"] = repr(exc)
    finally:
        native.set_ozaki(0)
    return out


def main():
    engines = [("fp64_dmma", None), ("int8_crt18_min256", (256, 18)), ("int8_crt16_min256", (256, 16)),
               ("int8_digits8_min256", (256, 8)), ("int8_crt16_min8192", (8192, 16))]
    rep = {}
    for name in T.CASES:
        rep[name] = {}
        for e in engines:
            rep[name][e[0]] = run(name, e)
            print(name, e[0], json.dumps(rep[name][e[0]]), flush=True)
    if len(sys.argv) > 1:
        json.dump(rep, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
