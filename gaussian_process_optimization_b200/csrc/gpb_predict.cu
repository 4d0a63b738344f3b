// Small kernels around the predictive GEMMs (north_star (c)): posterior mean, variance from the triangular product,
// the GPyOpt acquisition epilogue and the anchor-point top-k.
//
// Reference path being replaced: PosteriorExact._raw_predict (GPy/.../posterior.py:273-302), GPModel._predict / predict /
// predict_withGradients (GPyOpt/GPyOpt/models/gpmodel.py:95-142), get_quantiles (GPyOpt/util/general.py:113-128),
// AcquisitionEI (acquisitions/EI.py:32-51), AcquisitionLCB (LCB.py:31-46), AcquisitionBase sign/cost (base.py:33-50),
// anchor selection np.argsort(scores)[:k] (optimization/anchor_points_generator.py:58-63).
#include <float.h>

#include <algorithm>

#include "gpb_common.cuh"
#include "gpb_kernels.cuh"

namespace gpb {

__global__ void rowdot_kernel(const double *__restrict__ KxT, int ld, int n_c, int n, const double *__restrict__ alpha, int ld_alpha,
                              int p, double *__restrict__ mu) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= n_c) return;
  const double *row = KxT + (size_t)c * ld;
  for (int pp = 0; pp < p; ++pp) {
    const double *a = alpha + (size_t)pp * ld_alpha;
    double acc = 0.0;
    for (int j = lane; j < n; j += 32) acc = fma(row[j], a[j], acc);
    acc = warp_sum(acc);
    if (lane == 0) mu[(size_t)c * p + pp] = acc;
  }
}

int launch_rowdot(const double *KxT, int ld, int n_c, int n, const double *alpha, int ld_alpha, int p, double *mu, cudaStream_t s) {
  if (n_c == 0) return 0;
  rowdot_kernel<<<(n_c * 32 + 255) / 256, 256, 0, s>>>(KxT, ld, n_c, n, alpha, ld_alpha, p, mu);
  count_launch();
  GPB_CHECK_LAUNCH();
  return 0;
}

__global__ void var_from_vt_kernel(const double *__restrict__ Vt, int ld, int n_c, int n, double base, double *__restrict__ var) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= n_c) return;
  const double *row = Vt + (size_t)c * ld;
  double acc = 0.0;
  for (int j = lane; j < n; j += 32) acc = fma(row[j], row[j], acc);
  acc = warp_sum(acc);
  if (lane == 0) var[c] = base - acc;
}

int launch_var_from_vt(const double *Vt, int ld, int n_c, int n, double base, double *var, cudaStream_t s) {
  if (n_c == 0) return 0;
  var_from_vt_kernel<<<(n_c * 32 + 255) / 256, 256, 0, s>>>(Vt, ld, n_c, n, base, var);
  count_launch();
  GPB_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// A handful of candidates (the M = 1 .. 8 calls L-BFGS-B makes from every anchor point, optimizer.py:46-51, and estimate_L's
// calls, batch_local_penalization.py:56-58): one warp per candidate leaves 147 SMs idle and serialises N / 32 dependent
// iterations.  Here the training points are split over `splits` CTAs per candidate; one kernel forms the partial sums of
//     mu = Kx^T alpha,   sum_n Vt^2,   d mu / dx* = gradients_X(alpha^T),   d var / dx* = gradients_X(-2 Kx^T Ky^-1)
// (posterior.py:276,294; core/gp.py:431-434,450-453; stationary.py:354-364) and a second one adds the partials in a fixed order
// (bitwise reproducible run to run) and applies the scalings.  Slots per (candidate, split): [mu, vv, g1[DCAP], g2[DCAP]].
// ---------------------------------------------------------------------------------------------------------------------
constexpr int SKM_THREADS = 128;

template <int KIND, int DCAP>
__global__ void __launch_bounds__(SKM_THREADS) skinny_moments_partial_kernel(
    const double *__restrict__ KxT, const double *__restrict__ Vt, const double *__restrict__ Ut, int ld, int n, int chunk,
    const double *__restrict__ alpha, const double *__restrict__ XcT, int ldc, const double *__restrict__ XT, int ldx, int d,
    double variance, int want_g, double *__restrict__ part) {
  constexpr int K = 2 + 2 * DCAP;
  __shared__ double red[SKM_THREADS / 32][K];
  const int c = blockIdx.y, sp = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int j0 = sp * chunk, j1 = min(n, j0 + chunk);
  double acc[K];
#pragma unroll
  for (int i = 0; i < K; ++i) acc[i] = 0.0;
  double xc[DCAP];
#pragma unroll
  for (int q = 0; q < DCAP; ++q) xc[q] = (want_g && q < d) ? XcT[(size_t)q * ldc + c] : 0.0;
  for (int j = j0 + tid; j < j1; j += SKM_THREADS) {
    const double a = alpha[j];
    if (KxT) acc[0] = fma(KxT[(size_t)c * ld + j], a, acc[0]);
    if (Vt) {
      const double v = Vt[(size_t)c * ld + j];
      acc[1] = fma(v, v, acc[1]);
    }
    if (want_g) {
      double df[DCAP], r2 = 0.0;
#pragma unroll
      for (int q = 0; q < DCAP; ++q) {
        df[q] = (q < d) ? xc[q] - XT[(size_t)q * ldx + j] : 0.0;
        r2 = fma(df[q], df[q], r2);
      }
      double k, dk;
      cov_k_dk<KIND>(r2, variance, k, dk);
      const double w1 = dk * a;
      const double w2 = Ut ? dk * Ut[(size_t)c * ld + j] : 0.0;
#pragma unroll
      for (int q = 0; q < DCAP; ++q) {
        acc[2 + q] = fma(w1, df[q], acc[2 + q]);
        acc[2 + DCAP + q] = fma(w2, df[q], acc[2 + DCAP + q]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < K; ++i) {
    const double v = warp_sum(acc[i]);
    if (lane == 0) red[warp][i] = v;
  }
  __syncthreads();
  for (int i = tid; i < K; i += SKM_THREADS) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < SKM_THREADS / 32; ++w) v += red[w][i];
    part[((size_t)c * gridDim.x + sp) * K + i] = v;
  }
}

// grid (K, n_c): slot i of candidate c = sum over the splits, then scaled and stored where the one-warp kernels put it
template <int DCAP>
__global__ void __launch_bounds__(128) skinny_moments_reduce_kernel(const double *__restrict__ part, int splits, int d, double var_base,
                                                                    const double *__restrict__ inv_ls, double s1, double s2,
                                                                    double *__restrict__ mu, double *__restrict__ var,
                                                                    double *__restrict__ dmu, double *__restrict__ dvar) {
  constexpr int K = 2 + 2 * DCAP;
  __shared__ double scratch[32];
  const int i = blockIdx.x, c = blockIdx.y;
  double acc = 0.0;
  for (int sp = threadIdx.x; sp < splits; sp += 128) acc += part[((size_t)c * splits + sp) * K + i];
  acc = block_sum<128>(acc, scratch);
  if (threadIdx.x != 0) return;
  if (i == 0) {
    if (mu) mu[c] = acc;
  } else if (i == 1) {
    if (var) var[c] = var_base - acc;
  } else if (i < 2 + DCAP) {
    const int q = i - 2;
    if (dmu && q < d) dmu[(size_t)c * d + q] = s1 * inv_ls[q] * acc;
  } else {
    const int q = i - 2 - DCAP;
    if (dvar && q < d) dvar[(size_t)c * d + q] = s2 * inv_ls[q] * acc;
  }
}

size_t skinny_moments_part_doubles(int n_c, int np, int d) {
  const int dcap = d <= 4 ? 4 : d <= 8 ? 8 : d <= 16 ? 16 : 32;
  return (size_t)n_c * (np / SKM_THREADS) * (2 + 2 * dcap);
}

template <int KIND, int DCAP>
static int launch_skinny_moments_t(const double *KxT, const double *Vt, const double *Ut, int ld, int n_c, int n, const double *alpha,
                                   const double *XcT, int ldc, const double *XT, int ldx, int d, double variance,
                                   const double *inv_ls, double var_base, int want_g, double *part, double *mu, double *var,
                                   double *dmu, double *dvar, cudaStream_t s) {
  // about four CTAs per SM for ONE candidate, at least one 128-point chunk each.  The split of the training points must not depend
  // on n_c: a candidate's sums are then bit-identical whether it is evaluated alone or together with up to seven others, which is
  // what lets the host coalesce the M = 1 requests of concurrent L-BFGS-B runs into one call without changing any trajectory.
  const int max_splits = (n + SKM_THREADS - 1) / SKM_THREADS;
  int splits = std::max(1, std::min(max_splits, 148 * 4));
  const int chunk = ((n + splits - 1) / splits + SKM_THREADS - 1) / SKM_THREADS * SKM_THREADS;
  splits = (n + chunk - 1) / chunk;
  skinny_moments_partial_kernel<KIND, DCAP><<<dim3(splits, n_c), SKM_THREADS, 0, s>>>(KxT, Vt, Ut, ld, n, chunk, alpha, XcT, ldc, XT, ldx,
                                                                                    d, variance, want_g, part);
  GPB_CHECK_LAUNCH();
  skinny_moments_reduce_kernel<DCAP><<<dim3(2 + 2 * DCAP, n_c), 128, 0, s>>>(part, splits, d, var_base, inv_ls, 1.0, -2.0, mu, var, dmu,
                                                                            dvar);
  count_launch(2);
  GPB_CHECK_LAUNCH();
  return 0;
}

// mu (KxT != NULL), var = var_base - sum Vt^2 (Vt != NULL), dmu (want_g), dvar (want_g and Ut != NULL) for n_c <= 8 candidates
int launch_skinny_moments(int kind, const double *KxT, const double *Vt, const double *Ut, int ld, int n_c, int n, const double *alpha,
                          const double *XcT, int ldc, const double *XT, int ldx, int d, double variance, const double *inv_ls,
                          double var_base, int want_g, double *part, double *mu, double *var, double *dmu, double *dvar,
                          cudaStream_t s) {
  if (n_c == 0) return 0;
  GPB_REQUIRE(d <= 32, "skinny moments: input dimension %d > 32 takes the generic kernels (predict_block)", d);
#define GPB_SKM(K_, D_)                                                                                                          \
  return launch_skinny_moments_t<K_, D_>(KxT, Vt, Ut, ld, n_c, n, alpha, XcT, ldc, XT, ldx, d, variance, inv_ls, var_base, want_g, \
                                         part, mu, var, want_g ? dmu : nullptr, (want_g && Ut) ? dvar : nullptr, s)
  if (kind == GPB_KERN_RBF) {
    if (d <= 4) GPB_SKM(GPB_KERN_RBF, 4);
    if (d <= 8) GPB_SKM(GPB_KERN_RBF, 8);
    if (d <= 16) GPB_SKM(GPB_KERN_RBF, 16);
    GPB_SKM(GPB_KERN_RBF, 32);
  } else {
    if (d <= 4) GPB_SKM(GPB_KERN_MATERN52, 4);
    if (d <= 8) GPB_SKM(GPB_KERN_MATERN52, 8);
    if (d <= 16) GPB_SKM(GPB_KERN_MATERN52, 16);
    GPB_SKM(GPB_KERN_MATERN52, 32);
  }
#undef GPB_SKM
}

// One thread per candidate.  var already includes the likelihood variance (GP.predict include_likelihood=True).
__global__ void acq_epilogue_kernel(int acq, double par, double fmin, int n_c, int d, const double *__restrict__ mu,
                                    const double *__restrict__ var, const double *__restrict__ dmu, const double *__restrict__ dvar,
                                    double *__restrict__ f, double *__restrict__ df, double *__restrict__ mean_out,
                                    double *__restrict__ sd_out, double *__restrict__ dmdx_out, double *__restrict__ dsdx_out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_c) return;
  const double m = mu[c];
  double v = var[c];
  v = fmax(v, 1e-10);                       // gpmodel.py:99,137 np.clip(v, 1e-10, inf)
  double s = sqrt(v);                        // gpmodel.py:112,142
  const double ds_scale = 1.0 / (2.0 * s);   // gpmodel.py:140 dsdx = dvdx / (2 sqrt(v))
  double fa, c_m, c_s;                       // dacq = c_s * dsdx + c_m * dmdx
  if (acq == GPB_ACQ_EI) {
    if (s < 1e-10) s = 1e-10;                // general.py:121-124 (in place: the floored s is what EI multiplies)
    const double u = (fmin - m - par) / s;   // general.py:125
    const double phi = exp(-0.5 * u * u) / sqrt(2.0 * M_PI);
    const double Phi = 0.5 * erfc(-u / sqrt(2.0));
    fa = s * (u * Phi + phi);                // EI.py:39
    c_s = phi;                               // EI.py:50
    c_m = -Phi;
  } else {
    fa = -m + par * s;                       // LCB.py:36
    c_s = par;                               // LCB.py:45
    c_m = -1.0;
  }
  if (f) f[c] = -fa;                         // base.py:39,50 (cost == 1, indicator == 1)
  if (mean_out) mean_out[c] = m;
  if (sd_out) sd_out[c] = s;
  if (dmu) {
    for (int q = 0; q < d; ++q) {
      const double dm = dmu[(size_t)c * d + q];
      const double ds = dvar[(size_t)c * d + q] * ds_scale;
      if (df) df[(size_t)c * d + q] = -(c_s * ds + c_m * dm);
      if (dmdx_out) dmdx_out[(size_t)c * d + q] = dm;
      if (dsdx_out) dsdx_out[(size_t)c * d + q] = ds;
    }
  }
}

int launch_acq_epilogue(int acq, double par, double fmin, int n_c, int d, const double *mu, const double *var, const double *dmu,
                        const double *dvar, double *f, double *df, double *mean_out, double *sd_out, double *dmdx_out,
                        double *dsdx_out, cudaStream_t s) {
  if (n_c == 0) return 0;
  acq_epilogue_kernel<<<(n_c + 255) / 256, 256, 0, s>>>(acq, par, fmin, n_c, d, mu, var, dmu, dvar, f, df, mean_out, sd_out,
                                                        dmdx_out, dsdx_out);
  count_launch();
  GPB_CHECK_LAUNCH();
  return 0;
}

// ---- local penalisation (GPyOpt/GPyOpt/acquisitions/LP.py) -------------------------------------------------------------
// log Phi(z) without underflow: erfcx(x) = exp(x^2) erfc(x) keeps the left tail exact (scipy's norm.logcdf = log_ndtr).
__device__ __forceinline__ double log_ndtr(double z) {
  const double t = z * 0.70710678118654752440;
  if (z < 0.0) return log(0.5 * erfcx(-t)) - t * t;
  return log1p(-0.5 * erfc(t));
}

// One thread per candidate.  In: F = -acq and dF = -dacq as AcquisitionBase returns them (base.py:33-50).  Out:
//   f  = -T(acq) - sum_b log Phi((|x - x_b| - r_b) / s_b)                       LP.py:70-89 (_penalized_acquisition)
//   df = scale * dF - sum_b d_b   (d_b a scalar broadcast over the dimensions, exactly like LP.py:91-104,121-128)
// with T = log(acq + 1e-50), scale = 1 / acq (transform 'none') or T = log softplus(acq), scale = 1 / (softplus(acq) (1 + e^-acq))
// ('softplus', the default for LCB: LP.py:31-34).
__global__ void lp_epilogue_kernel(int n_c, int d, int nb, const double *__restrict__ Xc, const double *__restrict__ Xb,
                                   const double *__restrict__ r, const double *__restrict__ s, int transform,
                                   const double *__restrict__ F, const double *__restrict__ dF, double *__restrict__ f_out,
                                   double *__restrict__ df_out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_c) return;
  const double acqv = -F[c];
  double t, scale;
  if (transform == 1) {
    const double sp = log1p(exp(acqv));
    t = (acqv >= 40.0) ? log(acqv) : log(sp);
    scale = 1.0 / (sp * (1.0 + exp(-acqv)));
  } else {
    t = log(acqv + 1e-50);
    scale = 1.0 / acqv;
  }
  double hsum = 0.0, dsum = 0.0;
  for (int b = 0; b < nb; ++b) {
    double n2 = 0.0;
    for (int q = 0; q < d; ++q) {
      const double df = Xc[(size_t)c * d + q] - Xb[(size_t)b * d + q];
      n2 = fma(df, df, n2);
    }
    const double nm = sqrt(n2);
    const double z = (nm - r[b]) / s[b];
    hsum += log_ndtr(z);
    if (df_out) {
      const double hf = 0.5 * erfc(-z * 0.70710678118654752440);
      double dd = 1.0 / (s[b] * 2.50662827463100050242 * hf) * exp(-0.5 * z * z) / nm;
      if (hf < 1e-50) dd = 0.0;
      dsum += dd;
    }
  }
  f_out[c] = -t - hsum;
  if (df_out)
    for (int q = 0; q < d; ++q) df_out[(size_t)c * d + q] = scale * dF[(size_t)c * d + q] - dsum;
}

int launch_lp_epilogue(int n_c, int d, int nb, const double *Xc, const double *Xb, const double *r, const double *s, int transform,
                       const double *F, const double *dF, double *f_out, double *df_out, cudaStream_t st) {
  if (n_c == 0) return 0;
  lp_epilogue_kernel<<<(n_c + 127) / 128, 128, 0, st>>>(n_c, d, nb, Xc, Xb, r, s, transform, F, dF, f_out, df_out);
  count_launch();
  GPB_CHECK_LAUNCH();
  return 0;
}

// ---- top-k (k smallest, lexicographic on (value, global index)) -----------------------------------------------------
__device__ __forceinline__ bool lex_less(double av, long long ai, double bv, long long bi) {
  return (av < bv) || (av == bv && ai < bi);
}

__global__ void topk_init_kernel(double *vals, long long *idx, int k) {
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    vals[i] = DBL_MAX;
    idx[i] = LLONG_MAX;
  }
}

int launch_topk_init(double *vals, long long *idx, int k, cudaStream_t s) {
  topk_init_kernel<<<1, 64, 0, s>>>(vals, idx, k);
  count_launch();
  GPB_CHECK_LAUNCH();
  return 0;
}

constexpr int TOPK_MAX = 64;

// Single CTA: k selection rounds; each round takes the lexicographic minimum of {f[i]} U {previous state} that is strictly
// greater than the element selected in the previous round.  Deterministic.
__global__ void __launch_bounds__(1024) topk_update_kernel(const double *__restrict__ f, int n_c, long long index_base, double *vals,
                                                           long long *idx, int k) {
  __shared__ double sv[32];
  __shared__ long long si[32];
  __shared__ double old_v[TOPK_MAX], new_v[TOPK_MAX];
  __shared__ long long old_i[TOPK_MAX], new_i[TOPK_MAX];
  __shared__ double last_v;
  __shared__ long long last_i;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < k) {
    old_v[tid] = vals[tid];
    old_i[tid] = idx[tid];
  }
  if (tid == 0) {
    last_v = -DBL_MAX;
    last_i = -1;
  }
  __syncthreads();
  for (int r = 0; r < k; ++r) {
    const double lv = last_v;
    const long long li = last_i;
    const bool first = (r == 0);
    double bv = DBL_MAX;
    long long bi = LLONG_MAX;
    for (int i = tid; i < n_c; i += 1024) {
      const double v = f[i];
      const long long gi = index_base + i;
      if ((first || lex_less(lv, li, v, gi)) && lex_less(v, gi, bv, bi)) {
        bv = v;
        bi = gi;
      }
    }
    if (tid < k) {
      const double v = old_v[tid];
      const long long gi = old_i[tid];
      if ((first || lex_less(lv, li, v, gi)) && lex_less(v, gi, bv, bi)) {
        bv = v;
        bi = gi;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (lex_less(ov, oi, bv, bi)) {
        bv = ov;
        bi = oi;
      }
    }
    if (lane == 0) {
      sv[warp] = bv;
      si[warp] = bi;
    }
    __syncthreads();
    if (warp == 0) {
      bv = sv[lane];
      bi = si[lane];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (lex_less(ov, oi, bv, bi)) {
          bv = ov;
          bi = oi;
        }
      }
      if (lane == 0) {
        new_v[r] = bv;
        new_i[r] = bi;
        last_v = bv;
        last_i = bi;
      }
    }
    __syncthreads();
  }
  if (tid < k) {
    vals[tid] = new_v[tid];
    idx[tid] = new_i[tid];
  }
}

int launch_topk_update(const double *f, int n_c, long long index_base, double *vals, long long *idx, int k, cudaStream_t s) {
  GPB_REQUIRE(k >= 1 && k <= TOPK_MAX, "top-k: k must be in [1, %d]", TOPK_MAX);
  topk_update_kernel<<<1, 1024, 0, s>>>(f, n_c, index_base, vals, idx, k);
  count_launch();
  GPB_CHECK_LAUNCH();
  return 0;
}

// out[i] = [value, (double) global index, the candidate's d coordinates] for the k slots; a slot that never received a candidate
// (fewer than k finite scores) becomes [NaN, -1, NaN ...] -- the "empty" convention of sharded.merge_topk
__global__ void topk_pack_kernel(const double *__restrict__ vals, const long long *__restrict__ idx, const double *__restrict__ Xc, int d,
                                 int k, long long index_offset, double *__restrict__ out) {
  const int i = blockIdx.x;
  if (i >= k) return;
  const long long gi = idx[i];
  const bool empty = gi == LLONG_MAX;
  const double nan = __longlong_as_double(0x7ff8000000000000ll);
  double *row = out + (size_t)i * (d + 2);
  if (threadIdx.x == 0) {
    row[0] = empty ? nan : vals[i];
    row[1] = empty ? -1.0 : (double)gi;
  }
  for (int q = threadIdx.x; q < d; q += blockDim.x) row[2 + q] = empty ? nan : Xc[(size_t)(gi - index_offset) * d + q];
}

int launch_topk_pack(const double *vals, const long long *idx, const double *Xc_dev, int d, int k, long long index_offset, double *out,
                     cudaStream_t s) {
  topk_pack_kernel<<<k, 32, 0, s>>>(vals, idx, Xc_dev, d, k, index_offset, out);
  count_launch();
  GPB_CHECK_LAUNCH();
  return 0;
}

__global__ void min_kernel(const double *__restrict__ v, int n, double *out) {
  __shared__ double sv[32];
  double m = DBL_MAX;
  for (int i = threadIdx.x; i < n; i += 1024) m = fmin(m, v[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmin(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) sv[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = sv[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmin(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (threadIdx.x == 0) *out = m;
  }
}

// Componentwise backward error of the solve (Oettli-Prager): omega = max_i |Ky alpha - y|_i / (|Ky| |alpha| + |y|)_i, one warp per row
// of the full symmetric Ky (np x np), maximum taken with an integer atomic on the bit pattern (non-negative doubles order like
// integers; the maximum does not depend on the order of arrival).  *out must be zero before the launch.
__global__ void solve_residual_kernel(const double *__restrict__ Ky, int ld, int n, const double *__restrict__ alpha,
                                      const double *__restrict__ y, double *out) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= n) return;
  const double *row = Ky + (size_t)i * ld;
  double r = 0.0, sc = 0.0;
  for (int j = lane; j < n; j += 32) {
    const double t = row[j] * alpha[j];
    r += t;
    sc += fabs(t);
  }
  r = warp_sum(r);
  sc = warp_sum(sc);
  if (lane == 0) {
    const double yi = y[i];
    const double den = sc + fabs(yi);
    double om = den > 0.0 ? fabs(r - yi) / den : 0.0;
    if (!(om == om)) om = INFINITY;                        // NaN anywhere -> the check fails
    atomicMax(reinterpret_cast<unsigned long long *>(out), (unsigned long long)__double_as_longlong(om));
  }
}

int launch_solve_residual(const double *Ky, int ld, int n, const double *alpha, const double *y, double *out, cudaStream_t s) {
  GPB_CUDA(cudaMemsetAsync(out, 0, sizeof(double), s));
  solve_residual_kernel<<<(n * 32 + 255) / 256, 256, 0, s>>>(Ky, ld, n, alpha, y, out);
  count_launch();
  GPB_CHECK_LAUNCH();
  return 0;
}

int launch_min(const double *v, int n, double *out, cudaStream_t s) {
  min_kernel<<<1, 1024, 0, s>>>(v, n, out);
  count_launch();
  GPB_CHECK_LAUNCH();
  return 0;
}

}  // namespace gpb
