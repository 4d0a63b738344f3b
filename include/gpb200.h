/* gpb200.h -- C ABI of libgpb200.so: the B200 (sm_100a) exact-GP inner loop.
 *
 * Every entry point is `extern "C"`, takes plain pointers and sizes (no torch / C++ types) and returns an int status:
 *   0 = OK, <0 = bad argument / CUDA error (message via gpb_last_error()), >0 = LAPACK-style `info`
 *   (leading minor of that order is not positive definite).
 * Matrices are fp64, row-major (C-contiguous NumPy layout).  "dev" flags say whether a pointer is a device (1) or host (0)
 * pointer; host pointers are copied inside the call (the e2e path), device pointers are used in place.
 *
 * The reference (GPy 1.9.6 / GPyOpt 1.2.5) has no FFI on this path -- its plug-in boundary is Python duck typing
 * (SURVEY.md section 8b).  Each function below names the reference Python interface it replaces (file:line relative to
 * /root/reference); INTEGRATION.md shows the ctypes stub a maintainer would add on the reference side.
 */
#ifndef GPB200_H
#define GPB200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPB_KERN_RBF 0      /* GPy.kern.RBF        GPy/GPy/kern/src/rbf.py:50-54 */
#define GPB_KERN_MATERN52 1 /* GPy.kern.Matern52   GPy/GPy/kern/src/stationary.py:575-579 */
#define GPB_ACQ_EI 0        /* GPyOpt AcquisitionEI   GPyOpt/GPyOpt/acquisitions/EI.py:32-51 */
#define GPB_ACQ_LCB 1       /* GPyOpt AcquisitionLCB  GPyOpt/GPyOpt/acquisitions/LCB.py:31-46 */

#define GPB_ERR_CUDA (-1)   /* CUDA runtime error */
#define GPB_ERR_ARG (-2)    /* bad argument / wrong call order */
#define GPB_ERR_DOMAIN (-3) /* hyper-parameter outside its domain (NaN, inf, lengthscale <= 0, ...): the reference ends in
                               jitchol's LinAlgError for these (GPy/GPy/util/linalg.py:62-75) and so does the Python layer */

typedef struct gpb_model gpb_model; /* opaque: one GPRegression (data, hyper-parameters, posterior) resident on one GPU */

/* ---- library ------------------------------------------------------------------------------------------------------ */
int gpb_version(void);
const char *gpb_last_error(void);
int gpb_device_count(void);
/* Number of kernels this library has launched in the calling process (bench.py's `gpu_launches`). */
long long gpb_launch_count(void);

/* ---- Kern contract: GPy/GPy/kern/src/kern.py:119-202, stationary.py ------------------------------------------------ */
/* Stationary.K(X, X2=None)  (stationary.py:107-140, _scaled_dist :176-193, K_of_r rbf.py:50 / stationary.py:575).
 * X: n x d, X2: m x d or NULL (-> symmetric n x n, diagonal r forced to 0), lengthscale: nls = d (ARD) or 1 host doubles.
 * K out: n x m (ldk >= m). */
int gpb_kern_K(int kind, int d, int n, const double *X, int m, const double *X2, double variance,
               const double *lengthscale, int nls, double *K, int ldk, int dev, void *stream);

/* Stationary.update_gradients_full(dL_dK, X, X2=None)  (stationary.py:218-238, _inv_dist :251-258,
 * _lengthscale_grads_cython :263-269 -> stationary_cython.pyx:51-60).  dL_dK: n x m.  out (host): [dvariance, dlengthscale[nls]]. */
int gpb_kern_update_gradients_full(int kind, int d, int n, const double *X, int m, const double *X2, const double *dL_dK,
                                   int ld, double variance, const double *lengthscale, int nls, double *out, int dev,
                                   void *stream);

/* The reference's local "Gower" mixed-variable patch of Stationary.K (GPy/GPy/kern/src/stationary.py:61-65,116-135; enabled by
 * run.py:1207-1224 through GPyOpt/GPyOpt/models/gpmodel.py:58): K = prod_q k1d(r_q), r_q = |dx_q| / ranges[q] on continuous
 * dimensions and r_q = [x_q != x'_q] where discrete[q] != 0; the kernel's lengthscale is ignored by K.  The gradient methods
 * were NOT patched: update_gradients_full only swaps K in the variance term (:224), lengthscale and input gradients keep the
 * Euclidean distance -- reproduced as it is. */
int gpb_kern_K_gower(int kind, int d, int n, const double *X, int m, const double *X2, double variance, const int *discrete,
                     const double *ranges, double *K, int ldk, int dev, void *stream);
int gpb_kern_update_gradients_full_gower(int kind, int d, int n, const double *X, int m, const double *X2, const double *dL_dK,
                                         int ld, double variance, const double *lengthscale, int nls, const int *discrete,
                                         const double *ranges, double *out, int dev, void *stream);

/* Stationary.gradients_X(dL_dK, X, X2=None)  (stationary.py:271-278,354-364 -> stationary_utils.c:1-14 _grad_X).
 * out: n x d. */
int gpb_kern_gradients_X(int kind, int d, int n, const double *X, int m, const double *X2, const double *dL_dK, int ld,
                         double variance, const double *lengthscale, int nls, double *out, int dev, void *stream);

/* ---- util.linalg: GPy/GPy/util/linalg.py -------------------------------------------------------------------------- */
/* pdinv(A) (linalg.py:193-214) = jitchol's dpotrf (:56-60) + logdet (:208) + dpotri/symmetrify (:210-212).
 * A: n x n symmetric (lower triangle read).  Outputs (any may be NULL): L n x n lower (upper zeroed), Ai n x n symmetric,
 * Li n x n lower = L^-1, logdet (host double).  Returns info > 0 if a pivot is not positive (caller runs the jitter ladder,
 * linalg.py:62-75). */
int gpb_pdinv(int n, const double *A, int lda, double *L, double *Ai, double *Li, double *logdet, int dev, void *stream);
/* dpotrs(L, B) (linalg.py:116-125): solve (L L^T) X = B in place, B: n x nrhs row-major. */
int gpb_potrs(int n, const double *L, int ldl, double *B, int nrhs, int dev, void *stream);
/* dpotri(L) (linalg.py:127-145): A^-1 from the lower Cholesky factor, symmetrised. */
int gpb_potri(int n, const double *L, int ldl, double *Ai, int ldai, int dev, void *stream);

/* ---- GPRegression / ExactGaussianInference / PosteriorExact -------------------------------------------------------- */
/* Device workspace (bytes) a model of capacity n_cap points in d dims needs; cand_block = max candidates per predict block. */
size_t gpb_model_workspace_bytes(int n_cap, int d, int p, int cand_block);
/* GPRegression(X, Y, kernel, noise_var) (GPy/GPy/models/gp_regression.py:29-36, core/gp.py:38-110).
 * workspace: device memory owned by the caller (e.g. a torch tensor) of >= gpb_model_workspace_bytes, or NULL to let the
 * library cudaMalloc it.  ard: 1 -> d lengthscales, 0 -> one.  d <= 64, p <= 16. */
int gpb_model_create(gpb_model **out, int kind, int ard, int d, int p, int n_cap, int cand_block, void *workspace,
                     size_t workspace_bytes, void *stream);
int gpb_model_destroy(gpb_model *m);
/* GP.set_XY (core/gp.py:202-238).  X: n x d, Y: n x p. */
int gpb_model_set_data(gpb_model *m, int n, const double *X, const double *Y, int dev);
/* Matern52(..., Gower=True, space=...) (stationary.py:61-65): switch the model's covariance to the Gower product kernel.
 * discrete[d]: 1 for discrete dimensions; ranges[d]: domain width of the continuous ones (Design_space.lengthscales(),
 * GPyOpt/GPyOpt/core/task/space.py:351-362).  enable = 0 restores the stationary kernel. */
int gpb_model_set_gower(gpb_model *m, int enable, const int *discrete, const double *ranges);
/* parameter write: kern.variance, kern.lengthscale[nls], Gaussian_noise.variance (link order stationary.py:83, gp.py:108-109) */
int gpb_model_set_theta(gpb_model *m, double variance, const double *lengthscale, double noise);
/* GP.parameters_changed (core/gp.py:258-271) = ExactGaussianInference.inference (exact_gaussian_inference.py:37-74)
 * + Gaussian.update_gradients + Stationary.update_gradients_full.
 * extra_jitter is added to the diagonal on top of noise + 1e-8 (jitchol ladder, linalg.py:62-72).
 * out (host): [log_marginal, dL/dvariance, dL/dlengthscale[nls], dL/dnoise]  (want_grad = 0 -> only out[0]).
 * Returns 0 or info > 0 (not positive definite). */
int gpb_model_fit(gpb_model *m, int want_grad, double extra_jitter, double *out);
/* GP.set_XY (core/gp.py:202-238) when the new inputs extend the old ones and no hyper-parameter changed: appends b rows
 * Xnew (b x d) to the inputs, replaces the targets by Yall ((n + b) x p; GPyOpt re-normalises all of Y every step,
 * GPyOpt/GPyOpt/core/bo.py:246-247) and extends the resident factorisation by the new block rows in O(N^2 b) instead of
 * refactorising in O(N^3) (what GPModel.updateModel pays on every step, gpmodel.py:78-93 -> core/gp.py:258-271).
 * Needs a model fitted with extra_jitter = 0 for the current hyper-parameters; out as gpb_model_fit.  On info > 0 the
 * model must be refitted with gpb_model_fit. */
int gpb_model_append(gpb_model *m, int b, const double *Xnew, const double *Yall, int dev, int want_grad, double *out);
/* Posterior accessors (posterior.py:79-218): woodbury_chol L (n x n), woodbury_inv (n x n, symmetric), woodbury_vector
 * alpha (n x p), K (n x n, recomputed), dL_dK (n x n; exact_gaussian_inference.py:70).  dst may be host or device. */
int gpb_model_get(gpb_model *m, const char *what, double *dst, int ld, int dev);
/* GP.predict(Xnew, full_cov=False, include_likelihood) (core/gp.py:297-354; posterior.py:273-302; gaussian.py:102-110).
 * mu: mc x p, var: mc (no clipping, like PosteriorExact). */
int gpb_model_predict(gpb_model *m, int mc, const double *Xc, int include_likelihood, double *mu, double *var, int dev);
/* GP.predict(full_cov=True): cov mc x mc (posterior.py:281-284). */
int gpb_model_predict_full_cov(gpb_model *m, int mc, const double *Xc, int include_likelihood, double *mu, double *cov,
                               int dev);
/* GP.predictive_gradients(Xnew) (core/gp.py:407-454): dmu mc x d x p, dvar mc x d.  dvar = NULL skips the variance part
 * (estimate_L only uses the mean gradient, GPyOpt/GPyOpt/core/evaluators/batch_local_penalization.py:56-58). */
int gpb_model_predictive_gradients(gpb_model *m, int mc, const double *Xc, double *dmu, double *dvar, int dev);
/* GPModel.get_fmin (GPyOpt/GPyOpt/models/gpmodel.py:125-129): min posterior mean over the training inputs (host out). */
int gpb_model_fmin(gpb_model *m, double *fmin);

/* ---- GPyOpt acquisition: GPModel.predict(_withGradients) + get_quantiles + EI/LCB + AcquisitionBase sign ----------- */
/* acquisition_function(x) / acquisition_function_withGradients(x) (GPyOpt/GPyOpt/acquisitions/base.py:33-50) for an
 * unconstrained space and constant cost: f = -acq, df = -dacq.  par = jitter (EI) or exploration_weight (LCB).
 * Outputs (any may be NULL): f mc, df mc x d, mean mc, sd mc (clipped at 1e-10 like gpmodel.py:99), dmdx mc x d, dsdx mc x d. */
int gpb_model_acquisition(gpb_model *m, int acq, double par, double fmin, int mc, const double *Xc, double *f, double *df,
                          double *mean, double *sd, double *dmdx, double *dsdx, int dev);
/* AcquisitionLP.update_batches (GPyOpt/GPyOpt/acquisitions/LP.py:40-62): nb batch points Xb (nb x d) with their hammer-function
 * parameters r_x0, s_x0 (host arrays, copied to the device once); transform 0 = 'none', 1 = 'softplus'.  nb = 0 clears. */
int gpb_model_set_penalizers(gpb_model *m, int transform, int nb, const double *Xb, const double *r, const double *s);
/* AcquisitionLP.acquisition_function / acquisition_function_withGradients (LP.py:70-140): the log-transformed acquisition
 * penalised by the hammer functions set above.  f: mc, df: mc x d or NULL. */
int gpb_model_acquisition_lp(gpb_model *m, int acq, double par, double fmin, int mc, const double *Xc, double *f, double *df,
                             int dev);
/* Score mc candidates and return the k lowest f = -acq (anchor selection, anchor_points_generator.py:58-63, ties -> lowest
 * index).  idx are candidate indices + index_offset (global ids for a sharded candidate set).  vals/idx/pts are host. */
int gpb_model_acq_topk(gpb_model *m, int acq, double par, double fmin, int mc, const double *Xc, int dev, int k,
                       long long index_offset, double *vals, long long *idx, double *pts);

/* The same in ONE pass that also returns every candidate's f = -acq (mc; NULL to skip) and df = -dacq (mc x d; NULL to skip the
 * gradient solves): "score a candidate list with gradients and keep the k best" (run.py:1240-1253 at scale, BASELINE config 4). */
int gpb_model_acq_topk_full(gpb_model *m, int acq, double par, double fmin, int mc, const double *Xc, int dev, int k,
                            long long index_offset, double *vals, long long *idx, double *pts, double *f, double *df);

/* The same pass with everything left on the device and NO host synchronisation: candidates Xc_dev (mc x d), the k result rows
 * rows_dev (k x (d + 2): [f, global index as a double, the candidate's d coordinates]; a slot that never received a candidate --
 * fewer than k finite scores -- is [NaN, -1, NaN ...]), f_dev (mc, or NULL), df_dev (mc x d, or NULL to skip the gradient solves).
 * Results are ready in stream order on the stream the model was created with (for a model created on the legacy default stream --
 * which runs on a private non-blocking stream -- the legacy stream is made to wait for it), so that the all-gather of the per-shard anchors
 * (anchor_points_generator.py:58-63 on a sharded candidate set) can be issued behind it without a host round trip. */
int gpb_model_acq_topk_dev(gpb_model *m, int acq, double par, double fmin, int mc, const double *Xc_dev, int k,
                           long long index_offset, double *rows_dev, double *f_dev, double *df_dev);

/* ---- multi-GPU state distribution (broadcast of theta, alpha, L^-1 per model update) ------------------------------ */
/* Device address and length (doubles) of a resident array of the model, so that the caller can hand the collective library a view
 * of it (NCCL broadcast from the rank that fitted, instead of refitting on every rank: GPModel.updateModel, gpmodel.py:78-93).
 * what: "Li" = L^-1 (np x np, np = n rounded up to 128; what every predictive product reads), "alpha" (p x np),
 * "L" (np x np), "Wi" = Ky^-1 (np x np; computed first if the fit did not need it).  The arrays keep their padded layout. */
int gpb_model_state_ptr(gpb_model *m, const char *what, void **ptr, size_t *count);
/* After the regions above have been overwritten with another rank's fitted state: adopt its hyper-parameters (and the jitter its
 * fit needed) and mark the model fitted.  The data (gpb_model_set_data) must be the same on every rank.  have_mask: bit 0 = the
 * "L" region was transferred, bit 1 = "Wi" was (otherwise get("L") / append are refused and Wi is rebuilt from L^-1 on demand). */
int gpb_model_adopt_state(gpb_model *m, double variance, const double *lengthscale, double noise, double jitter, int have_mask);

/* ---- raw building blocks (tests, benchmarks, roofline measurements) ------------------------------------------------ */
/* C = alpha * op(A) op(B) + beta * C with the fp64 tensor-core (DMMA) engine; all of m, n multiples of 128, k of 16.
 * ta = 0: A is m x k row-major; 1: A is stored k x m.  tb = 0: B is stored n x k ("NT"); 1: B is stored k x n.
 * Device pointers only. */
int gpb_dgemm(int ta, int tb, int m, int n, int k, double alpha, const double *A, int lda, const double *B, int ldb,
              double beta, double *C, int ldc, void *stream);

/* Two-stream factorisation schedule of gpb_model_fit: blocks of at least min_n rows fork their T21 = L21 M11 product onto a
 * low-priority side stream underneath the critical path (default 512; 0 = single stream, used when timing kernels one by one). */
int gpb_set_overlap(int min_n);

/* Measurement hook for bench.py's roofline leg: when enabled, every DMMA GEMM launch is bracketed by CUDA events on its
 * stream.  collect() waits for them and returns the summed kernel time (ms), the executed tile flops and the launch count
 * since the last collect. */
int gpb_profile_gemm(int enable);
/* Tuning/testing knob: force the GEMM tile configuration (0 auto, 1 = 64x128 two CTAs/SM, 2 = 64x64 three CTAs/SM, 3 = 32x32,
 * 4 = 64x128 BK 32, 9 = 64x64 four CTAs/SM two stages); 100 + cfg only selects the configuration of the
 * large launches (>= 20 output blocks of 128x128) and leaves the small ones automatic. */
int gpb_gemm_config(int cfg);
int gpb_profile_gemm_collect(double *ms, double *flops, long long *launches);
/* Duration (ms) and executed flops of the last recorded GEMM launch (call before collect). */
int gpb_profile_gemm_last(double *ms, double *flops);

/* ---- experimental: fp64 products on the INT8 tensor cores (Ozaki scheme; csrc/gpb_ozaki.cu) ------------------------- */
/* C = alpha op(A) op(B)^T + beta C like gpb_dgemm, computed as `slices` (0 = the configured default, initially 8 = the maximum number
 * of digits; 7 is 25% faster and still within the parity bars of tests/; 10..18 = the modular mode described below, k <= 130944) balanced radix-256 digits per operand and
 * slices (slices + 1) / 2 exact int8 x int8 -> int32 products on tcgen05 (kind::i8, TMEM accumulators, TMA operands), recombined
 * in fp64.  tri_out / klo_mode / khi_mode: lower-tile output and per-tile k-ranges at 128 granularity (0 / 0 / 0 = plain product);
 * tri_a / tri_b: 0 = full operand, 1 = only the 128-blocks with k-block <= row-block are valid, 2 = k-block >= row-block.
 * m, n, k multiples of 128.  Device pointers only.  Replaces nothing in the reference by itself: it is an alternative engine for
 * the products inside pdinv (GPy/GPy/util/linalg.py:193-214). */
int gpb_ozaki_dgemm(int ta, int tb, int m, int n, int k, double alpha, const double *A, int lda, const double *B, int ldb,
                    double beta, double *C, int ldc, int tri_out, int klo_mode, int khi_mode, int tri_a, int tri_b, int slices,
                    void *stream);
/* Products of gpb_model_fit's factorisation with at least min_n rows go through the int8 engine (0 = off, the default; also
 * GPB_OZAKI_MIN_N / GPB_OZAKI_SLICES in the environment). */
int gpb_set_ozaki(int min_n, int slices);
/* Safety net of the engine inside gpb_model_fit.  Its products are accurate relative to the largest entry of an operand row, not
 * entry by entry, so every fit that used it measures the componentwise backward error of the solve Ky alpha = y against a Ky rebuilt
 * from the inputs; above the tolerance (2e-13, GPB_OZAKI_CHECK_TOL; 0 = no check) the evaluation is repeated on the fp64 DMMA engine
 * and the predictive products of that posterior stay there too.  engine_report: whether the current posterior came from the engine,
 * and the backward error measured (-1 if no check ran); fallback_count: fits of this process that were repeated. */
int gpb_model_engine_report(gpb_model *m, int *engine_used, double *residual);
long long gpb_ozaki_fallback_count(void);
/* Modular (CRT) mode of the same engine: slices in [10, 18] means "that many pairwise coprime moduli <= 256" instead of digits --
 * ONE int8 product per modulus (16 moduli carry 56 bits per operand at k = 16384, where 7 digits = 28 products carry 55), the
 * integer product rebuilt by the Chinese remainder theorem (csrc/gpb_crt.cuh).  gpb_ozaki_crt_bits: bits per operand for nmod moduli
 * and inner dimension k.  The two *_host_* entry points run the engine's integer arithmetic (residues of the scaled operand rows
 * [nmod][rows][k]; reconstruction of count integers from their [nmod][count] exact int32 residue-product sums) on the HOST: test
 * hooks that hold the device code bit-identical to oracle/ozaki_emulation.py without a GPU; nothing on the product path uses them. */
int gpb_ozaki_crt_bits(int nmod, long long k);
int gpb_ozaki_crt_host_residues(const double *A, int rows, int k, int nmod, int beta, signed char *planes, double *scale);
int gpb_ozaki_crt_host_combine(const int *sums, long long count, int nmod, double *X);

#ifdef __cplusplus
}
#endif
#endif /* GPB200_H */
