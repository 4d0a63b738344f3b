#!/usr/bin/env python
"""bench.py -- the driver's measurement contract for the exact-GP hot path.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Metric (BASELINE.json): NLL+grad evals/s (N=16384, D=16, RBF-ARD, fp64).  One step = one pass of the hot path
(K build -> Ky -> Cholesky + inverse -> alpha -> log-likelihood -> all D+2 gradients) over the synthetic data of SURVEY.md 8(d).

  value     evals/s with X, Y resident in HBM (CUDA events on the launching stream, max over ranks)
  e2e       the same through the host-buffer C-ABI call sequence (set_data: H2D of X, Y; set_theta; fit; D2H of D+3 doubles)
  roofline  the dominant kernel (DMMA GEMM engine): algorithmic N^3 flops per eval / its summed launch time per eval,
            against a cuBLAS DGEMM 8192^3 rate measured in the same process (MEASURED_PEAKS.json holds no fp64 figure)
  cpu_baseline  the CPU oracle (oracle/gp_oracle.py, NumPy/SciPy + the reference's own C helper) on the box's host cores
  aux       EI value+gradient candidates/s (config 4), candidates sharded over the ranks, top-5 all-gathered

Multi-GPU: the N x N factorisation stays on one GPU (north_star) -> NLL evals are independent replicas (weak scaling);
the acquisition shards its candidate set.  `--impl reference` times the CPU oracle (rank 0 only).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_TRAIN, DIM = 16384, 16
KIND = "rbf"
METRIC = "nll_grad_evals_per_s"
UNIT = "evals/s"
WORKLOAD = "GPRegression RBF-ARD N=16384 D=16 fp64 log_likelihood+gradients (SURVEY 8d headline)"


def synth(N, D, seed=1234):
    """SURVEY.md 8(d) generator (NumPy legacy RandomState so CPU and GPU legs share bits)."""
    rs = np.random.RandomState(seed)
    X = rs.uniform(0, 1, (N, D))
    w = rs.randn(D)
    Y = np.sin(X @ w)[:, None] + 0.05 * rs.randn(N, 1)
    Y = (Y - Y.mean()) / Y.std()
    ls = 0.5 + 0.5 * np.arange(D) / D
    return X, Y, ls


def algorithmic_flops(N, D):
    return float(N) ** 3 + float(N) ** 2 * (6 * D + 62)


# ----------------------------------------------------------------------------------------------------------------------
# CPU legs (oracle): the only place bench.py executes oracle/
# ----------------------------------------------------------------------------------------------------------------------
def cpu_eval_time(N, D, kind=KIND):
    """One oracle NLL+grad evaluation at (N, D); returns (total s, cubic LAPACK part s)."""
    from oracle import gp_oracle as O
    X, Y, ls = synth(N, D)
    t0 = time.perf_counter()
    Kmat = O.K(kind, X, None, 1.0, ls, True)
    Ky = Kmat.copy()
    Ky[np.diag_indices_from(Ky)] += 1e-2 + 1e-8
    t1 = time.perf_counter()
    Wi, LW, _, logdet = O.pdinv(Ky, with_Li=True)      # the reference's pdinv includes the unused dtrtri (linalg.py:209)
    alpha, _ = O.dpotrs(LW, Y, lower=1)
    t2 = time.perf_counter()
    dL_dK = 0.5 * (O.tdot(alpha) - Wi)
    O.update_gradients_full(kind, dL_dK, X, None, 1.0, ls, True, native=O.ref_native() is not None)
    t3 = time.perf_counter()
    return t3 - t0, t2 - t1


def cpu_baseline(sample_n=4096):
    import threadpoolctl
    cores = max([p.get("num_threads", 1) for p in threadpoolctl.threadpool_info()] + [1])
    cpu_eval_time(1024, DIM)  # warm up BLAS threads
    total, cubic = cpu_eval_time(sample_n, DIM)
    scale = N_TRAIN / sample_n
    est = cubic * scale ** 3 + (total - cubic) * scale ** 2
    from oracle import gp_oracle as O
    return {"value": 1.0 / est, "unit": UNIT, "cores": int(cores), "kind": "port",
            "sample": "one oracle eval at N=%d D=%d took %.2f s (LAPACK part %.2f s); extrapolated to N=%d as cubic x%d + "
                      "quadratic x%d = %.1f s/eval; lengthscale loop = reference C helper %s" %
                      (sample_n, DIM, total, cubic, N_TRAIN, scale ** 3, scale ** 2, est,
                       "(oracle/_ref)" if O.ref_native() is not None else "unavailable -> NumPy")}, est


def run_reference(args, rank):
    if rank != 0:
        return
    # each "step" = one bounded-sample oracle evaluation, reported in the metric's unit through the cubic/quadratic scaling
    sample_n = 2048
    ests = []
    for i in range(args.warmup + args.steps):
        total, cubic = cpu_eval_time(sample_n, DIM)
        scale = N_TRAIN / sample_n
        if i >= args.warmup:
            ests.append(cubic * scale ** 3 + (total - cubic) * scale ** 2)
    import threadpoolctl
    cores = max([p.get("num_threads", 1) for p in threadpoolctl.threadpool_info()] + [1])
    est = float(np.mean(ests))
    line = {"impl": "reference", "metric": METRIC, "value": 1.0 / est, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": est * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "note": "CPU oracle port of the reference path (the reference package cannot be "
                       "imported here: paramz is un-vendored); each step is an N=%d evaluation scaled to N=%d" % (sample_n, N_TRAIN)},
            "cpu_baseline": {"value": 1.0 / est, "unit": UNIT, "cores": int(cores), "kind": "port",
                             "sample": "N=%d D=%d oracle evals, cubic part x%d + quadratic part x%d" %
                                       (sample_n, DIM, (N_TRAIN // sample_n) ** 3, (N_TRAIN // sample_n) ** 2)},
            "e2e": {"value": 1.0 / est, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-ml): runs DURING the timed region
# ----------------------------------------------------------------------------------------------------------------------
class ClockSampler(object):
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nv = None

    def _loop(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.1)

    def start(self):
        if self._nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------------------------------------
def run_gpu(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from gaussian_process_optimization_b200 import native

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    X, Y, ls = synth(N_TRAIN, DIM)
    model = native.NativeModel(KIND, True, DIM, 1, n_cap=N_TRAIN, cand_block=2048)
    Xd, Yd = torch.from_numpy(X).cuda(), torch.from_numpy(Y).cuda()
    model.set_data(Xd, Yd)                 # device-resident inputs for `value`

    def theta(i):                          # every step evaluates a (slightly) different hyper-parameter vector, like L-BFGS-B
        return 1.0 + 1e-3 * (i % 7), ls * (1.0 + 1e-3 * (i % 5)), 1e-2

    def step_resident(i):
        v, l, nz = theta(i)
        model.set_theta(v, l, nz)
        info, logL, g = model.fit(True)
        assert info == 0 and np.isfinite(logL) and np.all(np.isfinite(g))
        return logL

    # e2e inputs live in pinned host memory (the C ABI takes plain host pointers; NumPy views of pinned torch tensors)
    Xp = torch.from_numpy(X).pin_memory()
    Yp = torch.from_numpy(Y).pin_memory()
    Xh, Yh = Xp.numpy(), Yp.numpy()

    def step_e2e(i):
        v, l, nz = theta(i)
        model.set_data(Xh, Yh)             # host buffers: H2D inside the timed region
        model.set_theta(v, l, nz)
        info, logL, g = model.fit(True)    # D2H of the D+3 results inside
        assert info == 0
        return logL

    # ---- fp64 peak: cuBLAS DGEMM 8192^3 in this process (burst, best of 5) ----
    n = 8192
    A = torch.randn(n, n, dtype=torch.float64, device="cuda")
    B = torch.randn(n, n, dtype=torch.float64, device="cuda")
    C = torch.empty(n, n, dtype=torch.float64, device="cuda")
    best = 1e30
    for i in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(A, B, out=C)
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            best = min(best, e0.elapsed_time(e1) * 1e-3)
    peak_tflops = 2.0 * n ** 3 / best / 1e12
    del A, B, C
    torch.cuda.empty_cache()

    # ---- warm-up ----
    for i in range(max(args.warmup, 3)):
        step_resident(i)
    step_e2e(0)

    # ---- timed: resident ----
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    launches0 = native.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step_resident(100 + i)
    e1.record()
    barrier()
    t_res = e0.elapsed_time(e1) * 1e-3
    launches = native.launch_count() - launches0

    # ---- timed: e2e ----
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step_e2e(200 + i)
    e1.record()
    barrier()
    t_e2e = e0.elapsed_time(e1) * 1e-3
    clocks = sampler.stop()

    # ---- roofline leg: per-launch events around the dominant kernel, same steps ----
    native.set_overlap(0)          # kernels one by one: per-launch events are only meaningful without concurrent kernels
    native.profile_gemm(1)
    for i in range(args.steps):
        step_resident(100 + i)
    last_ms, last_flops = native.profile_gemm_last()      # the last GEMM of an evaluation is Ky^-1 = M^T M, the largest launch
    gemm_ms, gemm_flops_exec, gemm_launches = native.profile_gemm_collect()
    native.profile_gemm(0)
    native.set_overlap(512)

    if world > 1:
        tt = torch.tensor([t_res, t_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_res, t_e2e = float(tt[0]), float(tt[1])

    # ---- aux: EI value+gradient candidates/s over a candidate set sharded across the ranks (config 4) ----
    aux = None
    try:
        aux = bench_acquisition(args, model, rank, world, barrier)
    except Exception as exc:  # report, never hide
        aux = {"error": repr(exc)}
    # ---- aux: the M = 1 value+gradient call L-BFGS-B makes from every anchor point (two HBM passes over the triangle of L^-1) ----
    aux_m1 = None
    if rank == 0:
        try:
            aux_m1 = bench_refinement_call(model)
        except Exception as exc:
            aux_m1 = {"error": repr(exc)}
    # ---- the other NLL+grad configurations of BASELINE.json (parity-test sizes; reported, not the headline) ----
    other = None
    if rank == 0:
        try:
            other = bench_other_configs(model)
        except Exception as exc:
            other = {"error": repr(exc)}

    # ---- aux: the same evaluation with the large products on the INT8 tensor cores (experimental engine, off by default) ----
    aux_int8 = None
    if rank == 0:
        try:
            aux_int8 = bench_int8_engine(args, model, theta, t_res / args.steps, peak_tflops)
        except Exception as exc:
            aux_int8 = {"error": repr(exc)}

    if rank == 0:
        value = world * args.steps / t_res
        gemm_s_per_eval = gemm_ms * 1e-3 / args.steps
        achieved = float(N_TRAIN) ** 3 / gemm_s_per_eval / 1e12
        traffic, traffic_note = None, None
        tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tp):
            try:
                tj = json.load(open(tp))
                traffic, traffic_note = tj.get("dram_bytes_per_launch"), tj.get("note")
            except Exception:
                traffic = None
        # the single largest launch (M^T M, lower tiles): algorithmic N^3 / 3 flops against its own event-timed duration
        dominant = {"launch": "gemm_dmma_kernel<COLK,COLK,64x64> Ky^-1 = M^T M (lower tiles), grid %d" % (2 * (N_TRAIN // 128) * (N_TRAIN // 128 + 1)),
                    "algorithmic_flops": float(N_TRAIN) ** 3 / 3.0, "ms": last_ms,
                    "achieved": (float(N_TRAIN) ** 3 / 3.0) / (last_ms * 1e-3) / 1e12 if last_ms > 0 else None,
                    "executed_tflops": last_flops / (last_ms * 1e-3) / 1e12 if last_ms > 0 else None,
                    "algorithmic_bytes": 3 * 8.0 * N_TRAIN * (N_TRAIN + 128) / 2, "traffic": traffic, "traffic_note": traffic_note}
        if dominant["achieved"]:
            dominant["frac"] = dominant["achieved"] / peak_tflops
        cpu, _ = cpu_baseline() if world == 1 else (None, None)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": t_res / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "parallelism": "replicas x%d (N x N factorisation stays on one GPU)" % world,
                       "l2": "inputs larger than L2 (three 2.1 GB fp64 matrices per eval vs 126 MB L2)",
                       "algorithmic_flops_per_eval": algorithmic_flops(N_TRAIN, DIM),
                       "algorithmic_tflops": algorithmic_flops(N_TRAIN, DIM) * args.steps / t_res / 1e12},
            "e2e": {"value": world * args.steps / t_e2e, "unit": UNIT,
                    "h2d_bytes_per_step": int(X.nbytes + Y.nbytes + (DIM + 2) * 8), "d2h_bytes_per_step": int((DIM + 3) * 8 + 4)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "gemm_dmma_kernel (fp64 DMMA.8x8x4 GEMM engine)", "achieved": achieved,
                         "peak": peak_tflops, "unit": "TFLOP/s", "frac": achieved / peak_tflops, "traffic": traffic,
                         "peak_source": "cuBLAS DGEMM 8192^3 measured in this run (burst, best of 5); MEASURED_PEAKS.json has no fp64 entry",
                         "launches_per_eval": gemm_launches / args.steps, "kernel_s_per_eval": gemm_s_per_eval,
                         "kernel_share_of_step": gemm_s_per_eval / (t_res / args.steps),
                         "executed_tflops": gemm_flops_exec / (gemm_ms * 1e-3) / 1e12, "dominant_launch": dominant},
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if aux is not None:
            line["aux"] = aux
        if aux_m1 is not None:
            line["aux_m1"] = aux_m1
        if other is not None:
            line["other_configs"] = other
        if aux_int8 is not None:
            line["aux_int8_engine"] = aux_int8
        print(json.dumps(line), flush=True)
    model.close()
    if world > 1:
        dist.destroy_process_group()


def bench_int8_engine(args, model, theta, dmma_s_per_eval, dgemm_peak_tflops):
    """NLL+grad at the headline size with every product of the factorisation that has >= 8192 rows on the int8 tensor cores
    (csrc/gpb_ozaki.cu: tcgen05 kind::i8, 7 balanced radix-256 digits per operand = 28 exact int8 products, fp64 recombination).
    Reported beside the headline, which stays on the fp64 DMMA engine: time per evaluation, agreement with the DMMA results on the
    same hyper-parameters, and the engine's own rate on an 8192^3 product against the library int8 GEMM measured in this run."""
    import torch
    from gaussian_process_optimization_b200 import native
    SLICES, MIN_N = 7, 8192
    v, l, nz = theta(3)
    model.set_theta(v, l, nz)
    info0, logL0, g0 = model.fit(True)
    # the EI value + gradient pass of config 4 on one shard, both engines on the same fitted hyper-parameters
    chunk = np.random.RandomState(4321).uniform(0, 1, (2 ** 15, DIM))
    shard = torch.from_numpy(np.ascontiguousarray(chunk)).cuda()
    fmin0 = model.fmin()
    vals0, idx0, pts0, f0, df0 = model.acq_topk_full("EI", 0.01, fmin0, shard, 5)
    f0, df0 = f0.cpu().numpy().copy(), df0.cpu().numpy().copy()
    native.set_ozaki(MIN_N, SLICES)
    try:
        for i in range(2):
            model.set_theta(*theta(i))
            model.fit(True)
        model.set_theta(v, l, nz)
        info1, logL1, g1 = model.fit(True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            model.set_theta(*theta(100 + i))
            info, logL, g = model.fit(True)
            assert info == 0 and np.isfinite(logL)
        e1.record()
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1) * 1e-3 / args.steps
        # the same with 8 digits per operand (36 products): indistinguishable from the fp64 engine, passes every parity test
        native.set_ozaki(MIN_N, 8)
        model.set_theta(v, l, nz)
        info8, logL8, g8 = model.fit(True)
        torch.cuda.synchronize()
        d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        d0.record()
        for i in range(max(2, args.steps // 2)):
            model.set_theta(*theta(100 + i))
            model.fit(True)
        d1.record()
        torch.cuda.synchronize()
        t8 = d0.elapsed_time(d1) * 1e-3 / max(2, args.steps // 2)
        eight = {"ms_per_eval": t8 * 1e3, "value": 1.0 / t8, "speedup_vs_dmma_engine": dmma_s_per_eval / t8,
                 "agreement_with_dmma_engine": {"logL_rel": abs(logL8 - logL0) / abs(logL0),
                                                "grad_rel_to_max": float(np.max(np.abs(g8 - g0)) / np.max(np.abs(g0)))}}
        # modular (CRT) mode of the same engine: 16 moduli = 16 int8 products per fp64 product (56 bits per operand at k = 16384,
        # where 7 digits = 28 products carry 55); the predictive products then use 18 moduli (62 bits) instead of 8 digits
        try:
            modular = bench_int8_modular(args, model, theta, (v, l, nz), dmma_s_per_eval, MIN_N, logL0, g0, shard, idx0, f0, df0)
        except Exception as exc:                       # the experimental block must never take the headline line down with it
            modular = {"error": repr(exc)}
        native.set_ozaki(MIN_N, SLICES)
        model.set_theta(v, l, nz)
        model.fit(True)
        fmin1 = model.fmin()
        model.acq_topk_full("EI", 0.01, fmin1, shard[:4096], 5)            # warm-up: cuts the digit planes of L^-1 once
        torch.cuda.synchronize()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        vals1, idx1, pts1, f1, df1 = model.acq_topk_full("EI", 0.01, fmin1, shard, 5)
        c1.record()
        torch.cuda.synchronize()
        t_acq = c0.elapsed_time(c1) * 1e-3
        acq = {"metric": "ei_value_gradient_candidates_per_s", "value": shard.shape[0] / t_acq, "unit": "candidates/s",
               "candidates": int(shard.shape[0]), "seconds": t_acq, "digits": 8,
               "algorithmic_tflops_fp64_equivalent": (2.0 * N_TRAIN ** 2 + N_TRAIN * (6 * DIM + 40)) * shard.shape[0] / t_acq / 1e12,
               "agreement_with_dmma_engine": {"same_top5": bool(list(idx0) == list(idx1)),
                                              "f_rel_to_max": float(np.max(np.abs(f1.cpu().numpy() - f0)) / np.max(np.abs(f0))),
                                              "df_rel_to_max": float(np.max(np.abs(df1.cpu().numpy() - df0)) / np.max(np.abs(df0)))}}
        # the engine alone on a square product, and the library int8 GEMM of the same shape as its roofline
        n = 8192
        A = torch.randn(n, n, dtype=torch.float64, device="cuda")
        B = torch.randn(n, n, dtype=torch.float64, device="cuda")
        C = torch.empty(n, n, dtype=torch.float64, device="cuda")
        best = 1e30
        for i in range(5):
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            native.ozaki_dgemm(0, 0, 1.0, A, B, 0.0, C, slices=SLICES)
            a1.record()
            torch.cuda.synchronize()
            if i >= 1:
                best = min(best, a0.elapsed_time(a1) * 1e-3)
        ref = A @ B.t()
        gemm_err = float((C - ref).abs().max() / ref.abs().max())
        del A, B, C, ref
        a8 = torch.randint(-127, 128, (n, n), dtype=torch.int8, device="cuda")
        b8 = torch.randint(-127, 128, (n, n), dtype=torch.int8, device="cuda")
        lib = 1e30
        for i in range(5):
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            torch._int_mm(a8, b8)
            a1.record()
            torch.cuda.synchronize()
            if i >= 1:
                lib = min(lib, a0.elapsed_time(a1) * 1e-3)
        del a8, b8
        pairs = SLICES * (SLICES + 1) // 2
        int8_tops = pairs * 2.0 * n ** 3 / best / 1e12
        lib_tops = 2.0 * n ** 3 / lib / 1e12
    finally:
        native.set_ozaki(0, SLICES)
    flops = algorithmic_flops(N_TRAIN, DIM)
    return {"metric": METRIC, "value": 1.0 / t, "unit": UNIT, "ms_per_eval": t * 1e3, "speedup_vs_dmma_engine": dmma_s_per_eval / t,
            "algorithmic_tflops_fp64_equivalent": flops / t / 1e12,
            "frac_of_dgemm_rate": flops / t / 1e12 / dgemm_peak_tflops,
            "engine": {"min_n": MIN_N, "digits": SLICES, "int8_products_per_fp64_product": pairs,
                       "products_on_the_engine": "the four products of the two top recursion levels and Ky^-1 = M^T M; everything below stays on DMMA"},
            "agreement_with_dmma_engine": {"logL_rel": abs(logL1 - logL0) / abs(logL0),
                                           "grad_rel_to_max": float(np.max(np.abs(g1 - g0)) / np.max(np.abs(g0))), "info": int(info1)},
            "with_8_digits": eight,
            "with_16_moduli": modular,
            "aux": acq,
            "roofline": {"bound": "tensor (int8, tcgen05 kind::i8)", "kernel": "ozaki_mma_kernel on an 8192^3 fp64-equivalent product (digit extraction included)",
                         "ms": best * 1e3, "effective_fp64_tflops": 2.0 * n ** 3 / best / 1e12, "achieved": int8_tops, "peak": lib_tops,
                         "unit": "TOP/s", "frac": int8_tops / lib_tops, "rel_err_vs_fp64_matmul": gemm_err,
                         "peak_source": "cuBLASLt int8 GEMM 8192^3 through torch._int_mm measured in this run (burst, best of 4); nominal dense int8 is 4500 TOP/s"},
            "note": "experimental, off by default (gpb_set_ozaki / GPB_OZAKI_MIN_N); the headline value, e2e and roofline above are the fp64 DMMA path"}


def bench_int8_modular(args, model, theta, theta0, dmma_s_per_eval, min_n, logL0, g0, shard, idx0, f0, df0, nmod=16):
    """The int8 engine in its modular mode (csrc/gpb_crt.cuh): evaluation time and agreement, the EI pass, one 8192^3 product."""
    import torch
    from gaussian_process_optimization_b200 import native
    native.set_ozaki(min_n, nmod)
    v, l, nz = theta0
    for i in range(2):
        model.set_theta(*theta(i))
        model.fit(True)
    model.set_theta(v, l, nz)
    info, logL, g = model.fit(True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        model.set_theta(*theta(100 + i))
        model.fit(True)
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) * 1e-3 / args.steps
    model.set_theta(v, l, nz)
    model.fit(True)
    fmin = model.fmin()
    model.acq_topk_full("EI", 0.01, fmin, shard[:4096], 5)                 # warm-up: cuts the residue planes of L^-1 once
    torch.cuda.synchronize()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    vals1, idx1, pts1, f1, df1 = model.acq_topk_full("EI", 0.01, fmin, shard, 5)
    c1.record()
    torch.cuda.synchronize()
    t_acq = c0.elapsed_time(c1) * 1e-3
    n = 8192
    A = torch.randn(n, n, dtype=torch.float64, device="cuda")
    B = torch.randn(n, n, dtype=torch.float64, device="cuda")
    C = torch.empty(n, n, dtype=torch.float64, device="cuda")
    best = 1e30
    for i in range(5):
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        native.ozaki_dgemm(0, 0, 1.0, A, B, 0.0, C, slices=nmod)
        a1.record()
        torch.cuda.synchronize()
        if i >= 1:
            best = min(best, a0.elapsed_time(a1) * 1e-3)
    ref = A @ B.t()
    gemm_err = float((C - ref).abs().max() / ref.abs().max())
    del A, B, C, ref
    flops = algorithmic_flops(N_TRAIN, DIM)
    return {"ms_per_eval": t * 1e3, "value": 1.0 / t, "speedup_vs_dmma_engine": dmma_s_per_eval / t,
            "algorithmic_tflops_fp64_equivalent": flops / t / 1e12,
            "engine": {"min_n": min_n, "moduli": nmod, "int8_products_per_fp64_product": nmod,
                       "bits_per_operand": native.ozaki_crt_bits(nmod, N_TRAIN)},
            "agreement_with_dmma_engine": {"logL_rel": abs(logL - logL0) / abs(logL0),
                                           "grad_rel_to_max": float(np.max(np.abs(g - g0)) / np.max(np.abs(g0))), "info": int(info)},
            "aux": {"metric": "ei_value_gradient_candidates_per_s", "value": shard.shape[0] / t_acq, "unit": "candidates/s",
                    "candidates": int(shard.shape[0]), "seconds": t_acq, "moduli": 18,
                    "agreement_with_dmma_engine": {"same_top5": bool(list(idx0) == list(idx1)),
                                                   "f_rel_to_max": float(np.max(np.abs(f1.cpu().numpy() - f0)) / np.max(np.abs(f0))),
                                                   "df_rel_to_max": float(np.max(np.abs(df1.cpu().numpy() - df0)) / np.max(np.abs(df0)))}},
            "product_8192": {"ms": best * 1e3, "effective_fp64_tflops": 2.0 * n ** 3 / best / 1e12, "rel_err_vs_fp64_matmul": gemm_err,
                             "kernels": "oz_absmax / oz_split (residues), ozaki_mma_kernel (16 accumulations), oz_crt_combine_kernel"}}


def bench_other_configs(model16k):
    """Device-resident NLL+grad time (best of 3 after one warm-up, CUDA events) for BASELINE.json configs 2 and 3."""
    import torch
    from gaussian_process_optimization_b200 import native
    out = {}
    for name, kind, N, D in (("config2_rbf_ard_n4096_d8", "rbf", 4096, 8), ("config3_mat52_ard_n16384_d16", "mat52", N_TRAIN, DIM)):
        X, Y, ls = synth(N, D)
        m = native.NativeModel(kind, True, D, 1, n_cap=N, cand_block=128)
        m.set_data(torch.from_numpy(X).cuda(), torch.from_numpy(Y).cuda())
        best = 1e30
        for i in range(4):
            m.set_theta(1.0 + 1e-3 * i, ls, 1e-2)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            info, logL, g = m.fit(True)
            e1.record()
            torch.cuda.synchronize()
            assert info == 0
            if i > 0:
                best = min(best, e0.elapsed_time(e1) * 1e-3)
        m.close()
        out[name] = {"ms_per_eval": best * 1e3, "evals_per_s": 1.0 / best,
                     "algorithmic_tflops": algorithmic_flops(N, D) / best / 1e12}
    return out


def bench_refinement_call(model):
    """EI value + gradient at ONE candidate (optimizer.py:46-51: what L-BFGS-B calls hundreds of times per BO step), host in /
    host out through the C ABI.  Bound: HBM -- M k* and M^T (M k*) each stream the lower triangle of M = L^-1 once."""
    import torch
    fmin = model.fmin()
    rs = np.random.RandomState(77)
    xs = rs.uniform(0, 1, (24, 1, DIM))
    for i in range(4):
        model.acquisition("EI", 0.01, fmin, xs[i], with_gradients=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(4, 24):
        model.acquisition("EI", 0.01, fmin, xs[i], with_gradients=True)
    torch.cuda.synchronize()
    t = (time.perf_counter() - t0) / 20
    nbytes = 2 * 8.0 * N_TRAIN * (N_TRAIN + 128) / 2          # two passes over the lower 128-blocks of M
    peak = None
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    out = {"metric": "ei_value_gradient_m1_calls_per_s", "value": 1.0 / t, "unit": "calls/s", "ms_per_call": t * 1e3, "model_N": N_TRAIN,
           "D": DIM, "roofline": {"bound": "hbm", "achieved": nbytes / t / 1e9, "unit": "GB/s", "peak": peak,
                                  "frac": (nbytes / t / 1e9 / peak) if peak else None,
                                  "algorithmic_bytes_per_call": nbytes,
                                  "note": "wall clock around the whole C-ABI call (H2D of x*, 9 launches, D2H of f and df), not the "
                                          "two streaming kernels alone; peak = MEASURED_PEAKS.json hbm_gbs"}}
    return out


def bench_acquisition(args, model, rank, world, barrier):
    """Config 4: EI value + gradient over a synthetic candidate set, sharded over the ranks; per-shard top-5 all-gathered."""
    import torch
    import torch.distributed as dist
    from gaussian_process_optimization_b200 import native
    per_rank = 2 ** 15
    chunk = np.random.RandomState(4321).uniform(0, 1, (2 ** 20, DIM))       # chunk c = RandomState(4321 + c)
    shard = torch.from_numpy(np.ascontiguousarray(chunk[rank * per_rank:(rank + 1) * per_rank])).cuda()
    model.set_theta(1.0, 0.5 + 0.5 * np.arange(DIM) / DIM, 1e-2)
    info, _, _ = model.fit(False)
    assert info == 0
    fmin = model.fmin()
    model.acq_topk_full("EI", 0.01, fmin, shard[:4096], 5)                   # warm-up
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    vals, idx, pts, f_all, df_all = model.acq_topk_full("EI", 0.01, fmin, shard, 5, index_offset=rank * per_rank)
    if world > 1:
        mine = torch.from_numpy(np.concatenate([vals, idx.astype(np.float64), pts.ravel()])).cuda()
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        allv = torch.stack(gathered).cpu().numpy()
        cand = sorted((allv[g, i], int(allv[g, 5 + i])) for g in range(world) for i in range(5))[:5]
    else:
        cand = sorted(zip(vals.tolist(), idx.tolist()))[:5]
    e1.record()
    barrier()
    t = e0.elapsed_time(e1) * 1e-3
    if world > 1:
        tt = torch.tensor([t], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t = float(tt[0])
    n_cand = per_rank * world
    fl = 2.0 * N_TRAIN ** 2 + N_TRAIN * (6 * DIM + 40)        # SURVEY 8(d) F_acq, value + gradient
    return {"metric": "ei_value_gradient_candidates_per_s", "value": n_cand / t, "unit": "candidates/s",
            "candidates": n_cand, "per_rank": per_rank, "model_N": N_TRAIN, "D": DIM, "seconds": t,
            "algorithmic_tflops": fl * n_cand / t / 1e12, "top5_global_idx": [c[1] for c in cand], "scaling": "weak",
            "note": "one pass: EI value + D-vector gradient for every candidate (written to HBM) and the running top-5; "
                    "per-shard top-5 all-gathered and merged inside the timed region"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gpb200", choices=["gpb200", "reference"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    run_gpu(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
