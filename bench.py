#!/usr/bin/env python
"""bench.py -- the driver's measurement contract for the exact-GP hot path.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Metric (BASELINE.json): "NLL+grad evals/s (N=16k, RBF-ARD, fp64); EI candidates/s at 1/2/4/8 GPU".  One step = one pass of the
hot path (K build -> Ky -> Cholesky + inverse -> alpha -> log-likelihood -> all D+2 gradients) over the synthetic data of
SURVEY.md 8(d).

  value     evals/s with X, Y resident in HBM (CUDA events on the launching stream, max over ranks)
  e2e       the same through the host-buffer C-ABI call sequence (set_data: H2D of X, Y; set_theta; fit; D2H of D+3 doubles)
  roofline  the dominant kernel (DMMA GEMM engine): algorithmic N^3 flops per eval / its summed launch time per eval,
            against a cuBLAS DGEMM 8192^3 rate measured in the same process (MEASURED_PEAKS.json holds no fp64 figure)
  cpu_baseline  the CPU oracle (oracle/gp_oracle.py, NumPy/SciPy + the reference's own C helper) on the box's host cores:
            ONE REAL evaluation at N=16384 (no extrapolation), all host threads pinned explicitly
  config.ei_*   the second half of the metric (BASELINE config 4): EI value + gradient + top-5 over a FIXED total of 2^20
            candidates split over the ranks (strong scaling), per-shard top-5 all-gathered on the device inside the timed region,
            the cost of getting the fitted state onto every rank (refit everywhere vs NCCL broadcast) reported beside it, and
            GPyOpt's real anchor-scoring call (1000 candidates, anchor_points_generator.py:87) at the same N

Multi-GPU: the N x N factorisation stays on one GPU (north_star) -> NLL evals are independent replicas (weak scaling);
the acquisition shards its candidate set (strong scaling, reported in config.ei_*).
`--impl reference` times the CPU oracle (rank 0 only): real evaluations at the full size, as many as fit its time budget.
"""
import os
import sys


def _pin_host_threads():
    """torch.distributed.run exports OMP_NUM_THREADS=1, which OpenBLAS and libgomp read when they are loaded: the CPU legs would
    run on one core.  Give them every core this process may use -- before NumPy / SciPy / the reference's C helper are imported."""
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    # GPU arm under torchrun (N > 1): no CPU leg runs there, leave the launcher's setting alone (N ranks x all cores would
    # oversubscribe the host)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and not any("reference" in a for a in sys.argv[1:]):
        return cores
    for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[var] = str(cores)
    return cores


HOST_CORES = _pin_host_threads()

import argparse   # noqa: E402
import json       # noqa: E402
import threading  # noqa: E402
import time       # noqa: E402

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_TRAIN, DIM = 16384, 16
KIND = "rbf"
METRIC = "nll_grad_evals_per_s"
UNIT = "evals/s"
WORKLOAD = "GPRegression RBF-ARD N=16384 D=16 fp64 log_likelihood+gradients (SURVEY 8d headline)"
EI_WORKLOAD = ("AcquisitionEI value+gradient+top-5 (BASELINE config 4): Matern52-ARD N=16384 D=16 noise 1e-6 model, 2^20 candidates "
               "(chunk 0 of SURVEY 8d) split over the ranks")
EI_TOTAL = 2 ** 20
EI_KIND, EI_NOISE = "mat52", 1e-6
N_ANCHOR_CANDIDATES = 1000        # GPyOpt's real anchor-scoring call (anchor_points_generator.py:87, random_design.py:56-77)


def base_config(world):
    """The keys both arms print (the driver compares the two config objects): workload names only, no measured values."""
    return {"workload": WORKLOAD, "N": N_TRAIN, "D": DIM, "kernel": "RBF-ARD", "noise": 1e-2,
            "ei_workload": EI_WORKLOAD, "ei_total_candidates": EI_TOTAL, "ei_model": "Matern52-ARD noise 1e-6",
            "anchor_candidates": N_ANCHOR_CANDIDATES}


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
def _blas_threads():
    """Pin every BLAS / OpenMP pool to HOST_CORES at run time as well (the environment alone is not enough when a pool was
    initialised earlier) and report what is actually in force."""
    import threadpoolctl
    try:
        threadpoolctl.threadpool_limits(limits=HOST_CORES)
    except Exception:
        pass
    return int(max([p.get("num_threads", 1) for p in threadpoolctl.threadpool_info()] + [1]))


def cpu_eval(N, D, kind=KIND, noise=1e-2, keep_state=False):
    """One oracle NLL+grad evaluation at (N, D), every step of GP.parameters_changed (core/gp.py:258-271) as the reference runs
    it: K, pdinv INCLUDING the dtrtri whose result exact inference never uses (linalg.py:209), dpotrs, dL_dK, the kernel gradient
    reductions with the reference's own compiled C loop.  -> (total s, cubic LAPACK part s, state or None)."""
    from oracle import gp_oracle as O
    X, Y, ls = synth(N, D)
    t0 = time.perf_counter()
    Kmat = O.K(kind, X, None, 1.0, ls, True)
    Ky = Kmat.copy()
    Ky[np.diag_indices_from(Ky)] += noise + 1e-8
    t1 = time.perf_counter()
    Wi, LW, _, logdet = O.pdinv(Ky, with_Li=True)
    alpha, _ = O.dpotrs(LW, Y, lower=1)
    t2 = time.perf_counter()
    dL_dK = 0.5 * (O.tdot(alpha) - Wi)
    O.update_gradients_full(kind, dL_dK, X, None, 1.0, ls, True, native=O.ref_native() is not None)
    t3 = time.perf_counter()
    state = None
    if keep_state:
        state = {"post": O.Posterior(LW, alpha, Kmat, Wi), "X": X, "ls": ls, "kind": kind, "noise": noise}
    return t3 - t0, t2 - t1, state


def cpu_ei_rate(state, m_sample=2048):
    """EI value + gradient on a bounded candidate sample with the oracle (GPModel.predict_withGradients + AcquisitionEI,
    gpmodel.py:131-142, EI.py:44-51), on the posterior of the evaluation just timed.  The reference recomputes get_fmin() -- a
    predict at all N training inputs -- inside every acquisition call (gpmodel.py:125-129, EI.py:36); it is timed separately and
    NOT charged to the per-candidate rate (which favours the CPU number)."""
    from oracle import gp_oracle as O
    post, X, ls, kind, noise = state["post"], state["X"], state["ls"], state["kind"], state["noise"]
    Xc = np.random.RandomState(4321).uniform(0, 1, (m_sample, X.shape[1]))
    t0 = time.perf_counter()
    fmin = O.gpmodel_get_fmin(kind, post, X, 1.0, ls, noise)
    t1 = time.perf_counter()
    m, s, dmdx, dsdx = O.gpmodel_predict_withGradients(kind, post, X, Xc, 1.0, ls, noise, native=O.ref_native() is not None)
    f, df = O.acq_EI(m, s, fmin, 0.01, dmdx, dsdx)
    t2 = time.perf_counter()
    return {"candidates_per_s": m_sample / (t2 - t1), "sample_candidates": m_sample, "seconds": t2 - t1, "get_fmin_seconds": t1 - t0}


def cpu_baseline():
    """cpu_baseline leg of the GPU arm (N = 1 only): ONE real oracle evaluation at the headline size ('single timed
    run for N=16384'), the small-N scaled estimate kept beside it as a cross-check.  GPB_BENCH_CPU=sample restores the bounded
    N=4096 sample (quick local runs)."""
    from oracle import gp_oracle as O
    cores = _blas_threads()
    helper = "(oracle/_ref)" if O.ref_native() is not None else "unavailable -> NumPy"
    cpu_eval(1024, DIM)                                   # spin the BLAS threads up
    total_s, cubic_s, _ = cpu_eval(4096, DIM)
    est = cubic_s * 64 + (total_s - cubic_s) * 16
    if os.environ.get("GPB_BENCH_CPU", "real") == "sample":
        return {"value": 1.0 / est, "unit": UNIT, "cores": cores, "kind": "port", "measured": False,
                "sample": "one oracle eval at N=4096 D=%d took %.2f s (LAPACK part %.2f s); EXTRAPOLATED to N=%d as cubic x64 + "
                          "quadratic x16 = %.1f s/eval; lengthscale loop = reference C helper %s" % (DIM, total_s, cubic_s, N_TRAIN, est, helper)}
    total, cubic, state = cpu_eval(N_TRAIN, DIM, keep_state=True)
    ei = None
    try:
        ei = cpu_ei_rate(state)
    except Exception as exc:
        ei = {"error": repr(exc)}
    return {"value": 1.0 / total, "unit": UNIT, "cores": cores, "kind": "port", "measured": True, "seconds_per_eval": total,
            "sample": "ONE REAL oracle eval at N=%d D=%d: %.1f s (LAPACK part %.1f s) on %d threads, no extrapolation; cross-check: "
                      "N=4096 eval %.2f s scaled (cubic x64 + quadratic x16) = %.1f s; lengthscale loop = reference C helper %s" %
                      (N_TRAIN, DIM, total, cubic, cores, total_s, est, helper),
            "ei": ei}


def run_reference(args, rank):
    """`--impl reference`: the reference's CPU path (oracle port; the reference package itself cannot be imported) at
    the FULL headline size on every host thread.  One evaluation takes minutes, so the arm runs as many real evaluations as fit
    its time budget (GPB_REF_BUDGET_S, default 420 s; at least one) and prints the count it actually ran as `steps`."""
    if rank != 0:
        return
    cores = _blas_threads()
    budget = float(os.environ.get("GPB_REF_BUDGET_S", "420"))
    cpu_eval(1024, DIM)                                   # spin the BLAS threads up (not a step)
    small_total, small_cubic, _ = cpu_eval(2048, DIM)
    est = small_cubic * 512 + (small_total - small_cubic) * 64
    times, state, t_start = [], None, time.perf_counter()
    while True:
        total, cubic, st = cpu_eval(N_TRAIN, DIM, keep_state=state is None)
        state = state or st
        times.append(total)
        elapsed = time.perf_counter() - t_start
        if len(times) >= args.steps or elapsed + 1.1 * total > budget:
            break
    per = float(np.mean(times))
    # the other half of the metric on the CPU: EI value + gradient on a bounded sample against a Matern52 posterior of the same size
    # would need a second 2-3 minute fit; the per-candidate cost does not depend on the kernel kind beyond the O(N D) covariance
    # row, so the RBF posterior of the evaluation above is used and the fact is stated
    try:
        ei = cpu_ei_rate(state)
    except Exception as exc:
        ei = {"error": repr(exc)}
    cfg = base_config(1)
    cfg.update({"ei_candidates_per_s": ei.get("candidates_per_s"), "ei_seconds": ei.get("seconds"),
                "ei_sample_candidates": ei.get("sample_candidates"), "ei_get_fmin_seconds": ei.get("get_fmin_seconds"),
                "ei_note": "CPU oracle, bounded sample, on the RBF posterior of the timed evaluation; get_fmin (recomputed by the "
                           "reference inside every acquisition call) timed separately and not charged",
                "note": "CPU oracle port of the reference path (the reference package cannot be imported here: paramz is "
                        "un-vendored); REAL evaluations at N=%d, %d run (requested %d) within a %.0f s budget; cross-check: N=2048 "
                        "eval %.2f s scaled (cubic x512 + quadratic x64) = %.1f s" % (N_TRAIN, len(times), args.steps, budget, small_total, est)})
    line = {"impl": "reference", "metric": METRIC, "value": 1.0 / per, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
            "warmup": 0, "steps_requested": args.steps, "warmup_requested": args.warmup,
            "ms_per_step": per * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": 1.0 / per, "unit": UNIT, "cores": cores, "kind": "port", "measured": True,
                             "sample": "%d real oracle evals at N=%d D=%d, %.1f s each (min %.1f, max %.1f) on %d threads" %
                                       (len(times), N_TRAIN, DIM, per, min(times), max(times), cores)},
            "e2e": {"value": 1.0 / per, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
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

    # ---- the second half of the metric: EI over 2^20 candidates, strong-scaled over the ranks (goes into `config`) ----
    ei = None
    try:
        ei = bench_ei_sharded(args, rank, world, barrier)
    except Exception as exc:  # report, never hide
        ei = {"ei_error": repr(exc)}
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
        cpu = cpu_baseline() if world == 1 else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": t_res / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": dict(base_config(world), **dict({
                "parallelism": "NLL+grad: replicas x%d (N x N factorisation stays on one GPU); EI: candidates sharded x%d" % (world, world),
                "l2": "inputs larger than L2 (three 2.1 GB fp64 matrices per eval vs 126 MB L2)",
                "algorithmic_flops_per_eval": algorithmic_flops(N_TRAIN, DIM),
                "algorithmic_tflops": algorithmic_flops(N_TRAIN, DIM) * args.steps / t_res / 1e12}, **(ei or {}))),
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
        if isinstance(aux_int8, dict) and isinstance(aux_int8.get("with_16_moduli"), dict) and "ms_per_eval" in aux_int8["with_16_moduli"]:
            mo = aux_int8["with_16_moduli"]
            line["config"]["int8_engine_16_moduli"] = {
                "note": "experimental engine, off by default; the headline stays on the fp64 DMMA path",
                "nll_grad_ms_per_eval": mo["ms_per_eval"], "nll_grad_evals_per_s": mo["value"],
                "ei_candidates_per_s_per_gpu_18_moduli": mo.get("aux", {}).get("value"),
                "agreement_logL_rel": mo.get("agreement_with_dmma_engine", {}).get("logL_rel"),
                "int8_tops_sustained": mo.get("roofline", {}).get("achieved"), "frac_of_nominal_int8_at_measured_clock": mo.get("roofline", {}).get("frac"),
                "sm_mhz_under_load": mo.get("roofline", {}).get("sm_mhz_under_load"),
                "engine_min_n": mo.get("engine", {}).get("min_n"),
                "nll_grad_ms_per_eval_fewer_moduli": {k: (w.get("ms_per_eval"), w.get("agreement_with_dmma_engine", {}).get("logL_rel"))
                                                      for k, w in mo.get("with_fewer_moduli", {}).items()}}
        if cpu is not None:
            line["cpu_baseline"] = cpu
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
            modular = bench_int8_modular(args, model, theta, (v, l, nz), dmma_s_per_eval, 4096, logL0, g0, shard, idx0, f0, df0)
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
    # the accuracy knob: fewer moduli = fewer int8 products.  Every evaluation is still checked (componentwise backward error of
    # Ky alpha = y, gpb_api.cu) and falls back to the fp64 engine when it fails -- at 12 moduli it does, and costs both evaluations.
    fewer = {}
    for nm in (14, 13):
        try:
            native.set_ozaki(min_n, nm)
            fb0 = native.ozaki_fallback_count()
            model.set_theta(*theta(0))
            model.fit(True)
            model.set_theta(v, l, nz)
            info_k, logL_k, g_k = model.fit(True)
            torch.cuda.synchronize()
            k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            k0.record()
            for i in range(args.steps):
                model.set_theta(*theta(100 + i))
                model.fit(True)
            k1.record()
            torch.cuda.synchronize()
            tk = k0.elapsed_time(k1) * 1e-3 / args.steps
            fewer["%d_moduli" % nm] = {"ms_per_eval": tk * 1e3, "value": 1.0 / tk, "speedup_vs_dmma_engine": dmma_s_per_eval / tk,
                                       "bits_per_operand": native.ozaki_crt_bits(nm, N_TRAIN),
                                       "fallbacks_to_fp64_engine": native.ozaki_fallback_count() - fb0,
                                       "agreement_with_dmma_engine": {"logL_rel": abs(logL_k - logL0) / abs(logL0),
                                                                      "grad_rel_to_max": float(np.max(np.abs(g_k - g0)) / np.max(np.abs(g0))),
                                                                      "info": int(info_k)}}
        except Exception as exc:
            fewer["%d_moduli" % nm] = {"error": repr(exc)}
    native.set_ozaki(min_n, nmod)
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
    # the engine's int8 rate against NOMINAL dense int8 (4500 TOP/s at the maximum SM clock) scaled to the clock it actually ran at:
    # a sustained second of products with the clock sampled underneath (the library int8 GEMM used as the denominator in round 1
    # moved 25% from run to run)
    sampler = ClockSampler(torch.cuda.current_device())
    sampler.start()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(20, int(1.0 / best))
    s0.record()
    for i in range(reps):
        native.ozaki_dgemm(0, 0, 1.0, A, B, 0.0, C, slices=nmod)
    s1.record()
    torch.cuda.synchronize()
    clk = sampler.stop()
    t_sus = s0.elapsed_time(s1) * 1e-3 / reps
    nominal = 4500.0 * (clk["sm_mhz"] / clk["sm_max_mhz"]) if clk.get("sm_mhz") and clk.get("sm_max_mhz") else None
    int8_tops = nmod * 2.0 * n ** 3 / t_sus / 1e12
    roof = {"bound": "tensor (int8, tcgen05 kind::i8)", "kernel": "one 8192^3 fp64-equivalent product = residue extraction + %d int8 "
            "products (ozaki_mma_kernel) + CRT reconstruction, sustained for %.1f s" % (nmod, reps * t_sus),
            "ms_per_product_sustained": t_sus * 1e3, "achieved": int8_tops, "unit": "TOP/s", "nominal_peak_at_max_clock": 4500.0,
            "sm_mhz_under_load": clk.get("sm_mhz"), "sm_max_mhz": clk.get("sm_max_mhz"), "throttle_reasons": clk.get("reasons"),
            "peak": nominal, "frac": (int8_tops / nominal) if nominal else None,
            "peak_source": "nominal dense int8 4500 TOP/s x (median SM clock during this loop / maximum SM clock)"}
    del A, B, C, ref
    flops = algorithmic_flops(N_TRAIN, DIM)
    return {"ms_per_eval": t * 1e3, "value": 1.0 / t, "speedup_vs_dmma_engine": dmma_s_per_eval / t, "roofline": roof,
            "algorithmic_tflops_fp64_equivalent": flops / t / 1e12,
            "engine": {"min_n": min_n, "moduli": nmod, "int8_products_per_fp64_product": nmod,
                       "bits_per_operand": native.ozaki_crt_bits(nmod, N_TRAIN)},
            "with_fewer_moduli": fewer,
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
    """EI value + gradient at M = 1, 5 and 8 candidates per call (optimizer.py:46-51: what L-BFGS-B calls hundreds of times per BO
    step; the host's LockstepEvaluator coalesces the anchors' requests into one M <= 8 call), host in / host out through the C ABI.
    Bound: HBM -- M k* and M^T (M k*) each stream the lower triangle of M = L^-1 once, whatever M <= 8 is."""
    import torch
    fmin = model.fmin()
    rs = np.random.RandomState(77)
    nbytes = 2 * 8.0 * N_TRAIN * (N_TRAIN + 128) / 2          # two passes over the lower 128-blocks of M
    peak = None
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    out = {}
    for mc in (1, 5, 8):
        xs = rs.uniform(0, 1, (44, mc, DIM))
        for i in range(4):
            model.acquisition("EI", 0.01, fmin, xs[i], with_gradients=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(4, 44):
            model.acquisition("EI", 0.01, fmin, xs[i], with_gradients=True)
        torch.cuda.synchronize()
        t = (time.perf_counter() - t0) / 40
        out["m%d" % mc] = {"ms_per_call": t * 1e3, "us_per_candidate": t * 1e6 / mc, "gbs": nbytes / t / 1e9,
                           "frac": (nbytes / t / 1e9 / peak) if peak else None}
    t1 = out["m1"]["ms_per_call"] * 1e-3
    return {"metric": "ei_value_gradient_m1_calls_per_s", "value": 1.0 / t1, "unit": "calls/s", "ms_per_call": t1 * 1e3, "model_N": N_TRAIN,
            "D": DIM, "batched": out,
            "roofline": {"bound": "hbm", "achieved": nbytes / t1 / 1e9, "unit": "GB/s", "peak": peak,
                         "frac": (nbytes / t1 / 1e9 / peak) if peak else None,
                         "algorithmic_bytes_per_call": nbytes,
                         "note": "wall clock around the whole C-ABI call (staged H2D of x*, the fused cooperative kernel "
                                 "skinny_fused_kernel + the acquisition epilogue, staged D2H of f and df, one synchronisation), not the "
                                 "kernel alone; peak = MEASURED_PEAKS.json hbm_gbs"}}


def bench_ei_sharded(args, rank, world, barrier):
    """The second half of BASELINE.json's metric, the one path that shards (north_star (c); run.py:1240-1253 /
    anchor_points_generator.py:85-98 scaled up): EI value + D-vector gradient for every candidate (written to HBM) and the global
    top-5, over a FIXED total of 2^20 candidates split into contiguous ranges over the ranks -> strong scaling.  The per-shard
    top-5 rows stay on the device, are all-gathered there and merged; all of that is inside the timed region (CUDA events, max
    over ranks).  Model = BASELINE config 4: Matern52-ARD, N=16384, D=16, noise 1e-6 (the exact_feval level).

    Also measured, outside that region: the cost of getting the fitted state onto every rank (every rank refits vs rank 0 fits and
    NCCL-broadcasts theta, alpha, L^-1), and GPyOpt's real anchor-scoring call (1000 uniform candidates, value only)."""
    import torch
    import torch.distributed as dist
    from gaussian_process_optimization_b200 import native, sharded
    X, Y, ls = synth(N_TRAIN, DIM)
    model = native.NativeModel(EI_KIND, True, DIM, 1, n_cap=N_TRAIN, cand_block=2048)
    model.set_data(torch.from_numpy(X).cuda(), torch.from_numpy(Y).cuda())
    model.set_theta(1.0, ls, EI_NOISE)

    def timed(fn, reps=1):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            out = fn()
        e1.record()
        barrier()
        t = e0.elapsed_time(e1) * 1e-3 / reps
        if world > 1:
            tt = torch.tensor([t], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t = float(tt[0])
        return t, out

    # ---- state distribution: (a) every rank refits; (b) rank 0 fits, the others receive theta / alpha / L^-1 over NCCL ----
    def refit():
        model.set_theta(1.0, ls * (1 + 1e-9), EI_NOISE)      # a changed hyper-parameter forces a real refit
        model.set_theta(1.0, ls, EI_NOISE)
        info, logL, _ = model.fit(False)
        assert info == 0
        return logL
    refit()
    t_refit, logL = timed(refit)
    t_bcast, state_ok = None, None
    if world > 1:
        model.broadcast_state(src=0)                         # warm-up (NCCL channel set-up)
        t_bcast, _ = timed(lambda: model.broadcast_state(src=0))
        # the adopted state must be the fitted one: compare alpha with what this rank computed itself before the broadcast
        a_mine = model.state_tensor("alpha").clone()
        refit()
        state_ok = bool(torch.equal(a_mine, model.state_tensor("alpha")))
        model.broadcast_state(src=0)
    fmin = model.fmin()

    # ---- strong scaling: 2^20 candidates in total ----
    lo, hi = sharded.divide_candidates(EI_TOTAL, rank, world)
    chunk = np.random.RandomState(4321).uniform(0, 1, (EI_TOTAL, DIM))          # chunk 0
    shard = torch.from_numpy(np.ascontiguousarray(chunk[lo:hi])).cuda()
    rows = torch.empty((5, DIM + 2), dtype=torch.float64, device="cuda")
    model.acq_topk_dev("EI", 0.01, fmin, shard[:4096], 5, index_offset=lo, with_gradients=True, rows=rows)   # warm-up
    torch.cuda.synchronize()

    def score():
        r, f, df = model.acq_topk_dev("EI", 0.01, fmin, shard, 5, index_offset=lo, with_gradients=True, rows=rows)
        return sharded.all_gather_topk_device(r, 5)
    t_ei, (tv, ti, tp) = timed(score)

    # ---- parity against the reference-generated fixture: top-5 of the first 2^16 rows (tests/golden/make_golden_fullsize.py) ----
    golden_ok = None
    gpath = os.path.join(ROOT, "tests", "golden", "fullsize", "config4_mat52_ard_n16384_d16_exact.npz")
    if rank == 0 and os.path.exists(gpath):
        z = np.load(gpath)
        pre = torch.from_numpy(np.ascontiguousarray(chunk[:2 ** 16])).cuda()
        v5, i5, _ = model.acq_topk("EI", 0.01, fmin, pre, 5)
        golden_ok = bool(np.array_equal(i5, z["ei_top5_idx"]) and np.allclose(v5, z["ei_top5_val"], rtol=1e-7, atol=0))
        del pre

    # ---- GPyOpt's real anchor scoring: 1000 uniform candidates, value only, top-5 (anchor_points_generator.py:58-63,87) ----
    anchors = np.random.RandomState(4322).uniform(0, 1, (N_ANCHOR_CANDIDATES, DIM))
    alo, ahi = sharded.divide_candidates(N_ANCHOR_CANDIDATES, rank, world)
    ash = torch.from_numpy(np.ascontiguousarray(anchors[alo:ahi])).cuda()

    def score_anchors():
        r, _, _ = model.acq_topk_dev("EI", 0.01, fmin, ash, min(5, ahi - alo), index_offset=alo, rows=rows[:min(5, ahi - alo)])
        return sharded.all_gather_topk_device(r, 5)
    score_anchors()
    t_anchor, (av, ai, ap) = timed(score_anchors, reps=10)
    del shard, chunk
    model.close()
    fl = 2.0 * N_TRAIN ** 2 + N_TRAIN * (6 * DIM + 40)        # SURVEY 8(d) F_acq, value + gradient
    return {"ei_candidates_per_s": EI_TOTAL / t_ei, "ei_seconds": t_ei, "ei_candidates_per_rank": hi - lo, "ei_scaling": "strong",
            "ei_algorithmic_tflops": fl * EI_TOTAL / t_ei / 1e12, "ei_top5_idx": [int(i) for i in ti],
            "ei_top5_val": [float(v) for v in tv], "ei_top5_prefix65536_matches_reference_golden": golden_ok,
            "ei_timed_region": "per-shard EI value+gradient+top-5 on device-resident candidates, device all-gather of the top-5 rows, merge",
            "ei_state_refit_every_rank_ms": t_refit * 1e3, "ei_state_broadcast_ms": None if t_bcast is None else t_bcast * 1e3,
            "ei_state_broadcast_bytes": int(8 * (N_TRAIN * N_TRAIN + N_TRAIN + 4 + DIM)) if world > 1 else 0,
            "ei_state_broadcast_equals_local_fit": state_ok,
            "anchor_call_ms": t_anchor * 1e3, "anchor_top5_idx": [int(i) for i in ai],
            "anchor_note": "GPyOpt's own anchor-scoring call (1000 uniform candidates, value only) split over the ranks: where sharding stops paying"}


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
