// Stationary-kernel construction K(X, X') and the hyper-parameter / input gradient reductions (north_star (a)).
//
// Replaces GPy/GPy/kern/src/stationary.py: _unscaled_dist/_scaled_dist (:155-193), K (:107-140), update_gradients_full
// (:218-238) with _inv_dist (:251-258) and the serial Cython loop lengthscale_grads (stationary_cython.pyx:51-60), and
// gradients_X (:271-278,354-364 -> stationary_utils.c:1-14); RBF.K_of_r/dK_dr (rbf.py:50-54), Matern52 (:575-579).
//
// Layout: inputs are pre-scaled once per parameter write into a dimension-major array XsT[q][i] = X[i][q] / l_q
// (row stride ldx >= number of points, zero padded), so that a 64-point tile of one dimension is one coalesced 512 B line
// and lands in shared memory in the order the register-tiled pair loops read it.  r^2 is the direct sum of squared
// differences (exactly 0 on the diagonal, never negative) instead of the reference's |x|^2+|y|^2-2xy expansion.
#include <cooperative_groups.h>
#include <cuda_pipeline.h>

#include "gpb_common.cuh"
#include "gpb_kernels.cuh"

namespace cg = cooperative_groups;

namespace gpb {

// XsT[q][i] = X[i][q] / ls[q]  (i < n), 0 for n <= i < ldx.   ls_dev: d doubles (already broadcast for the isotropic case)
__global__ void scale_transpose_kernel(const double *__restrict__ X, int n, int d, const double *__restrict__ ls, double *__restrict__ XsT,
                                       int ldx) {
  const size_t total = (size_t)d * ldx;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int q = (int)(e / ldx), i = (int)(e - (size_t)q * ldx);
    // |x / l| is clamped at 1e150: when an optimiser step drives a lengthscale towards 0 the scaled coordinates would overflow
    // to inf and produce inf - inf = NaN distances and 0 * inf = NaN gradient terms; clamped, coincident points keep r = 0,
    // distinct ones get r^2 ~ 1e300 -> k = 0 exactly, and every product below stays finite (no guards in the inner loops)
    double v = (i < n) ? X[(size_t)i * d + q] / ls[q] : 0.0;
    v = v > 1e150 ? 1e150 : (v < -1e150 ? -1e150 : v);
    XsT[e] = v;
  }
}

int launch_scale_transpose(const double *X, int n, int d, const double *ls_dev, double *XsT, int ldx, cudaStream_t s, int /*tag*/) {
  const size_t total = (size_t)d * ldx;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  scale_transpose_kernel<<<blocks, 256, 0, s>>>(X, n, d, ls_dev, XsT, ldx);
  count_launch();
  GPB_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// K tile kernel: one CTA = one 64 x 64 tile, 256 threads, thread (ty, tx) owns rows ty + 16a, cols tx + 16b (a, b < 4):
// for fixed (a, b) a warp writes two 128-byte row segments (fully coalesced).  The 4 x 4 register tile keeps the kernel at
// ~64 registers (4+ CTAs per SM) and its unrolled body inside the instruction cache -- the first version (8 x 8 per thread,
// 1 CTA per SM) stalled on instruction fetch and exposed latency (ncu: profiles/r1c_kgrad_full.md).
//   mode 0 (rect):  out[i][j] = k(r_ij) for i < n_rows, j < n_cols (nothing else is written)
//   mode 1 (Ky, padded): out is np x np;  i, j < n: k + (i == j) * diag_add;  otherwise identity
//   mode 2 (rect, zero padded): out is rows_pad x cols_pad; k inside n_rows x n_cols, 0 outside
//   mode 3 = mode 1 restricted to the 128-blocks on and below the diagonal (input of the factorisation)
// ---------------------------------------------------------------------------------------------------------------------
// GOWER = 1: the reference's local "Gower" patch (stationary.py:116-135) -- a product of one-dimensional kernels,
// r_q = |dx_q| / range_q on continuous dimensions (XaT / XbT then hold x_q / range_q) and r_q = [x_q != x'_q] on discrete ones
// (gflag[q] != 0; coordinates unscaled).  Every factor carries the variance: `variance` is then variance^d (vpow).
//   Matern52: prod_q (1 + s5 r_q + 5/3 r_q^2) * exp(-s5 sum_q r_q)        RBF: exp(-1/2 sum_q r_q^2)
// -- one exp per pair instead of d.
template <int KIND, int GOWER>
__global__ void __launch_bounds__(256, 4) kmat_kernel(const double *__restrict__ XaT, int lda, const double *__restrict__ XbT, int ldb, int d,
                                                      int n_rows, int n_cols, double variance, double diag_add, int mode,
                                                      double *__restrict__ out, int ldo, const double *__restrict__ gflag,
                                                      int row_blk0, const double *__restrict__ theta) {
  extern __shared__ double sm[];
  double *xa = sm;              // [d][64]
  double *xb = sm + d * KTILE;  // [d][64]
  if (theta) {                  // CUDA-graph replays: the hyper-parameters live in device memory, not in the baked arguments
    variance = theta[0];
    diag_add = theta[1];
  }
  const int row0 = (blockIdx.y + row_blk0) * KTILE, col0 = blockIdx.x * KTILE;   // row_blk0 > 0: only the rows from there on
  if (mode == 3) {
    // Ky for the factorisation: nothing reads the 128-blocks strictly above the diagonal (gpb_chol.cu only touches lower
    // blocks and complete diagonal blocks), so half of the exp() work and of the HBM writes is skipped
    if ((col0 >> 7) > (row0 >> 7)) return;
    mode = 1;
  }
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  for (int e = tid; e < d * KTILE; e += 256) {
    const int q = e >> 6, i = e & 63;
    xa[e] = XaT[(size_t)q * lda + row0 + i];
    xb[e] = XbT[(size_t)q * ldb + col0 + i];
  }
  __syncthreads();
  double r2[4][4], pr[GOWER ? 4 : 1][GOWER ? 4 : 1];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      r2[a][b] = 0.0;
      if (GOWER) pr[a][b] = 1.0;
    }
#pragma unroll 2
  for (int q = 0; q < d; ++q) {
    double va[4], vb[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) va[a] = xa[q * KTILE + ty + 16 * a];
#pragma unroll
    for (int b = 0; b < 4; ++b) vb[b] = xb[q * KTILE + tx + 16 * b];
    const bool disc = GOWER ? (gflag[q] != 0.0) : false;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const double df = va[a] - vb[b];
        if (!GOWER) {
          r2[a][b] = fma(df, df, r2[a][b]);
        } else {
          const double r = disc ? (df != 0.0 ? 1.0 : 0.0) : fabs(df);
          gower_accumulate<KIND>(r, r2[a][b], pr[a][b]);
        }
      }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int i = row0 + ty + 16 * a;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int j = col0 + tx + 16 * b;
      const bool inside = (i < n_rows) && (j < n_cols);
      double v;
      if (inside) {
        v = GOWER ? gower_value<KIND>(r2[a][b], pr[GOWER ? a : 0][GOWER ? b : 0], variance) : cov_k<KIND>(r2[a][b], variance);
        if (mode == 1 && i == j) v += diag_add;
      } else {
        v = (mode == 1 && i == j) ? 1.0 : 0.0;
      }
      if (inside || mode != 0) out[(size_t)i * ldo + j] = v;
    }
  }
}

int launch_kmat(int kind, const double *XaT, int lda, const double *XbT, int ldb, int d, int n_rows, int n_cols,
                double variance, double diag_add, int mode, double *out, int ldo, int rows_pad, int cols_pad,
                cudaStream_t s, const double *gflag, int row_start, const double *theta) {
  const size_t smem = (size_t)2 * d * KTILE * sizeof(double);
  GPB_REQUIRE(smem <= 200 * 1024, "input dimension %d too large", d);
  GPB_REQUIRE(rows_pad % KTILE == 0 && cols_pad % KTILE == 0, "kmat: padded sizes must be multiples of %d", KTILE);
  GPB_REQUIRE(row_start % KTILE == 0 && row_start >= 0 && row_start <= rows_pad, "kmat: bad row_start %d", row_start);
  dim3 grid(cols_pad / KTILE, (rows_pad - row_start) / KTILE);
  const int row_blk0 = row_start / KTILE;
  if (grid.x == 0 || grid.y == 0) return 0;
#define GPB_KMAT(K_, G_)                                                                                                  \
  do {                                                                                                                    \
    if (smem > 48 * 1024)                                                                                                 \
      GPB_CUDA(cudaFuncSetAttribute(kmat_kernel<K_, G_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));         \
    kmat_kernel<K_, G_><<<grid, 256, smem, s>>>(XaT, lda, XbT, ldb, d, n_rows, n_cols, variance, diag_add, mode, out, ldo, \
                                                gflag, row_blk0, theta);                                                  \
  } while (0)
  if (kind == GPB_KERN_RBF) {
    if (gflag) GPB_KMAT(GPB_KERN_RBF, 1); else GPB_KMAT(GPB_KERN_RBF, 0);
  } else {
    if (gflag) GPB_KMAT(GPB_KERN_MATERN52, 1); else GPB_KMAT(GPB_KERN_MATERN52, 0);
  }
#undef GPB_KMAT
  count_launch();
  GPB_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// Hyper-parameter gradient reduction.  Per 64 x 64 tile: recompute r^2, k, k'/r from the X tiles, form G on the fly and
// accumulate
//     part[tile][0]     = sum w K_ij G_ij                      (-> d/dvariance after / variance, stationary.py:224)
//     part[tile][1]     = sum_{i == j} G_ii                    (-> d/dnoise = tr(dL_dK), exact_gaussian_inference.py:72)
//     part[tile][2 + q] = sum w (k'/r)_ij G_ij (xs_iq - xs_jq)^2   (-> -(.)/l_q, stationary.py:227-235,260-269)
// FUSED = 1:  G = 0.5 (sum_p alpha_ip alpha_jp - P Wi_ij) (exact_gaussian_inference.py:70) is never materialised; only the
//             lower tiles of Wi are read (w = 2 off the diagonal tiles, the diagonal tiles are complete).
// FUSED = 0:  G = dL_dK[i][j] (n_rows x n_cols, any matrix) -- the Kern.update_gradients_full(dL_dK, X, X2) contract.
// Reductions: fixed-order shuffles -> per-warp shared slots -> per-tile partial -> a second fixed-order kernel; no atomics,
// so the result is bitwise reproducible (L-BFGS-B trajectories depend on it).
// ---------------------------------------------------------------------------------------------------------------------
template <int KIND, int FUSED>
__global__ void __launch_bounds__(256, 3) kgrad_kernel(const double *__restrict__ XaT, int lda, const double *__restrict__ XbT, int ldb, int d,
                                                       int n_rows, int n_cols, double variance, const double *__restrict__ G, int ldg,
                                                       const double *__restrict__ alpha, int ld_alpha, int p_out, int tiles_x,
                                                       double *__restrict__ part, const double *__restrict__ theta) {
  extern __shared__ double sm[];
  if (theta) variance = theta[0];    // CUDA-graph replays (see kmat_kernel)
  double *xa = sm;                   // [d][64]
  double *xb = sm + d * KTILE;       // [d][64]
  double *wacc = xb + d * KTILE;     // [8 warps][d + 2]
  int tr, tc;
  if (FUSED) {
    const int t = blockIdx.x;
    tr = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
    while ((tr + 1) * (tr + 2) / 2 <= t) ++tr;
    while (tr * (tr + 1) / 2 > t) --tr;
    tc = t - tr * (tr + 1) / 2;
  } else {
    tr = blockIdx.x / tiles_x;
    tc = blockIdx.x - tr * tiles_x;
  }
  const int row0 = tr * KTILE, col0 = tc * KTILE;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4, warp = tid >> 5, lane = tid & 31;
  for (int e = tid; e < d * KTILE; e += 256) {
    const int q = e >> 6, i = e & 63;
    xa[e] = XaT[(size_t)q * lda + row0 + i];
    xb[e] = XbT[(size_t)q * ldb + col0 + i];
  }
  // G (or Wi) entries of this thread: issued before the distance loop so that the HBM latency hides behind it
  double gw[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int i = row0 + ty + 16 * a;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int j = col0 + tx + 16 * b;
      gw[a][b] = (i < n_rows && j < n_cols) ? G[(size_t)i * ldg + j] : 0.0;
    }
  }
  __syncthreads();
  double r2[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) r2[a][b] = 0.0;
#pragma unroll 2
  for (int q = 0; q < d; ++q) {
    double va[4], vb[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) va[a] = xa[q * KTILE + ty + 16 * a];
#pragma unroll
    for (int b = 0; b < 4; ++b) vb[b] = xb[q * KTILE + tx + 16 * b];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const double df = va[a] - vb[b];
        r2[a][b] = fma(df, df, r2[a][b]);
      }
  }
  if (FUSED) {
    // gw <- dL_dK = 0.5 (alpha alpha^T - P Wi)
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int i = row0 + ty + 16 * a;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int j = col0 + tx + 16 * b;
        if (i < n_rows && j < n_cols) {
          double aa = 0.0;
          for (int p = 0; p < p_out; ++p) aa = fma(alpha[(size_t)p * ld_alpha + i], alpha[(size_t)p * ld_alpha + j], aa);
          gw[a][b] = 0.5 * (aa - (double)p_out * gw[a][b]);
        }
      }
    }
  }
  // r2[a][b] is overwritten by the pair weight  w * (k'/r) * G
  const double wt = (FUSED && tr != tc) ? 2.0 : 1.0;
  double acc_var = 0.0, acc_tr = 0.0;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int i = row0 + ty + 16 * a;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int j = col0 + tx + 16 * b;
      double wgt = 0.0;
      if (i < n_rows && j < n_cols) {
        const double g = gw[a][b];
        if (FUSED && i == j) acc_tr += g;
        double k, dk;
        cov_k_dk<KIND>(r2[a][b], variance, k, dk);
        acc_var = fma(wt * k, g, acc_var);
        wgt = wt * dk * g;
      }
      r2[a][b] = wgt;
    }
  }
  acc_var = warp_sum(acc_var);
  acc_tr = warp_sum(acc_tr);
  if (lane == 0) {
    wacc[warp * (d + 2) + 0] = acc_var;
    wacc[warp * (d + 2) + 1] = acc_tr;
  }
  for (int q = 0; q < d; ++q) {
    double va[4], vb[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) va[a] = xa[q * KTILE + ty + 16 * a];
#pragma unroll
    for (int b = 0; b < 4; ++b) vb[b] = xb[q * KTILE + tx + 16 * b];
    double acc4[4] = {0.0, 0.0, 0.0, 0.0};   // one chain per row block: four independent dependency chains
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const double df = va[a] - vb[b];
        acc4[a] = fma(r2[a][b], df * df, acc4[a]);
      }
    double acc = (acc4[0] + acc4[1]) + (acc4[2] + acc4[3]);
    acc = warp_sum(acc);
    if (lane == 0) wacc[warp * (d + 2) + 2 + q] = acc;
  }
  __syncthreads();
  for (int c = tid; c < d + 2; c += 256) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += wacc[w * (d + 2) + c];
    part[(size_t)blockIdx.x * (d + 2) + c] = s;
  }
}

// out[c] = sum_t part[t][c], fixed order (one block per column).
__global__ void colsum_kernel(const double *__restrict__ part, int ntiles, int ncols, double *__restrict__ out) {
  __shared__ double scratch[32];
  const int c = blockIdx.x;
  double acc = 0.0;
  for (int t = threadIdx.x; t < ntiles; t += 256) acc += part[(size_t)t * ncols + c];
  acc = block_sum<256>(acc, scratch);
  if (threadIdx.x == 0) out[c] = acc;
}

// Gower variant of the variance-gradient term only:  part[tile] = sum w K_gower,ij G_ij  (stationary.py:224 with the patched K;
// the lengthscale terms keep the Euclidean distance in the reference, so kgrad_kernel still produces them).
template <int KIND, int FUSED>
__global__ void __launch_bounds__(256, 3) kvar_gower_kernel(const double *__restrict__ XaT, int lda, const double *__restrict__ XbT, int ldb,
                                                            int d, int n_rows, int n_cols, double vpow, const double *__restrict__ gflag,
                                                            const double *__restrict__ G, int ldg, const double *__restrict__ alpha,
                                                            int ld_alpha, int p_out, int tiles_x, double *__restrict__ part) {
  extern __shared__ double sm[];
  double *xa = sm, *xb = sm + d * KTILE, *wacc = xb + d * KTILE;
  int tr, tc;
  if (FUSED) {
    const int t = blockIdx.x;
    tr = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
    while ((tr + 1) * (tr + 2) / 2 <= t) ++tr;
    while (tr * (tr + 1) / 2 > t) --tr;
    tc = t - tr * (tr + 1) / 2;
  } else {
    tr = blockIdx.x / tiles_x;
    tc = blockIdx.x - tr * tiles_x;
  }
  const int row0 = tr * KTILE, col0 = tc * KTILE;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4, warp = tid >> 5, lane = tid & 31;
  for (int e = tid; e < d * KTILE; e += 256) {
    const int q = e >> 6, i = e & 63;
    xa[e] = XaT[(size_t)q * lda + row0 + i];
    xb[e] = XbT[(size_t)q * ldb + col0 + i];
  }
  __syncthreads();
  double sacc[4][4], pr[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      sacc[a][b] = 0.0;
      pr[a][b] = 1.0;
    }
  for (int q = 0; q < d; ++q) {
    const bool disc = gflag[q] != 0.0;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const double df = xa[q * KTILE + ty + 16 * a] - xb[q * KTILE + tx + 16 * b];
        const double r = disc ? (df != 0.0 ? 1.0 : 0.0) : fabs(df);
        gower_accumulate<KIND>(r, sacc[a][b], pr[a][b]);
      }
  }
  const double wt = (FUSED && tr != tc) ? 2.0 : 1.0;
  double acc = 0.0;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int i = row0 + ty + 16 * a;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int j = col0 + tx + 16 * b;
      if (i < n_rows && j < n_cols) {
        double g = G[(size_t)i * ldg + j];
        if (FUSED) {
          double aa = 0.0;
          for (int p = 0; p < p_out; ++p) aa = fma(alpha[(size_t)p * ld_alpha + i], alpha[(size_t)p * ld_alpha + j], aa);
          g = 0.5 * (aa - (double)p_out * g);
        }
        acc = fma(wt * gower_value<KIND>(sacc[a][b], pr[a][b], vpow), g, acc);
      }
    }
  }
  acc = warp_sum(acc);
  if (lane == 0) wacc[warp] = acc;
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += wacc[w];
    part[blockIdx.x] = s;
  }
}

// out_dev[0] = sum K_gower . G  (part: scratch of `tiles` doubles)
int launch_kvar_gower(int kind, int fused, const double *XaT, int lda, const double *XbT, int ldb, int d, int n_rows, int n_cols,
                      double vpow, const double *gflag, const double *G, int ldg, const double *alpha, int ld_alpha, int p_out,
                      double *part, double *out_dev, cudaStream_t s) {
  const size_t smem = (size_t)(2 * d * KTILE + 8) * sizeof(double);
  GPB_REQUIRE(smem <= 200 * 1024, "input dimension %d too large", d);
  const int tr = (n_rows + KTILE - 1) / KTILE, tc = (n_cols + KTILE - 1) / KTILE;
  const int tiles = fused ? tr * (tr + 1) / 2 : tr * tc;
  if (tiles == 0) return 0;
#define GPB_KV(K_, F_)                                                                                                     \
  do {                                                                                                                     \
    if (smem > 48 * 1024)                                                                                                  \
      GPB_CUDA(cudaFuncSetAttribute(kvar_gower_kernel<K_, F_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
    kvar_gower_kernel<K_, F_><<<tiles, 256, smem, s>>>(XaT, lda, XbT, ldb, d, n_rows, n_cols, vpow, gflag, G, ldg, alpha,    \
                                                       ld_alpha, p_out, tc, part);                                          \
  } while (0)
  if (kind == GPB_KERN_RBF) {
    if (fused) GPB_KV(GPB_KERN_RBF, 1); else GPB_KV(GPB_KERN_RBF, 0);
  } else {
    if (fused) GPB_KV(GPB_KERN_MATERN52, 1); else GPB_KV(GPB_KERN_MATERN52, 0);
  }
#undef GPB_KV
  GPB_CHECK_LAUNCH();
  colsum_kernel<<<1, 256, 0, s>>>(part, tiles, 1, out_dev);
  count_launch(2);
  GPB_CHECK_LAUNCH();
  return 0;
}

int launch_kgrad(int kind, int fused, const double *XaT, int lda, const double *XbT, int ldb, int d, int n_rows, int n_cols,
                 double variance, const double *G, int ldg, const double *alpha, int ld_alpha, int p_out, double *part,
                 double *out_dev, cudaStream_t s, const double *theta) {
  const size_t smem = (size_t)(2 * d * KTILE + 8 * (d + 2)) * sizeof(double);
  GPB_REQUIRE(smem <= 200 * 1024, "input dimension %d too large", d);
  const int tr = (n_rows + KTILE - 1) / KTILE, tc = (n_cols + KTILE - 1) / KTILE;
  const int tiles = fused ? tr * (tr + 1) / 2 : tr * tc;
  if (tiles == 0) return 0;
#define GPB_KGRAD(K_, F_)                                                                                                 \
  do {                                                                                                                    \
    if (smem > 48 * 1024)                                                                                                 \
      GPB_CUDA(cudaFuncSetAttribute(kgrad_kernel<K_, F_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));        \
    kgrad_kernel<K_, F_><<<tiles, 256, smem, s>>>(XaT, lda, XbT, ldb, d, n_rows, n_cols, variance, G, ldg, alpha, ld_alpha, \
                                                  p_out, tc, part, theta);                                                \
  } while (0)
  if (kind == GPB_KERN_RBF) {
    if (fused) GPB_KGRAD(GPB_KERN_RBF, 1); else GPB_KGRAD(GPB_KERN_RBF, 0);
  } else {
    if (fused) GPB_KGRAD(GPB_KERN_MATERN52, 1); else GPB_KGRAD(GPB_KERN_MATERN52, 0);
  }
#undef GPB_KGRAD
  GPB_CHECK_LAUNCH();
  colsum_kernel<<<d + 2, 256, 0, s>>>(part, tiles, d + 2, out_dev);
  count_launch(2);
  GPB_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// Input gradients:  out1[c][q] = (s1 / l_q) sum_n (k'/r)_cn g1_cn (xs_cq - xs_nq)       [and the same with g2, s2]
// with g_cn = G[c * ldg + n] (ldg == 0 broadcasts one row: the dL_dK = alpha^T case of core/gp.py:431-434)
// (+ G[n * ldg + c] when add_t: the `tmp + tmp.T` of stationary.py:359-361 when X2 is None).
// gradx_kernel (up to 8 rows): GX_WPC warps per output row c; lanes stride over n (coalesced in the dimension-major layout);
// per-lane accumulators for all dimensions live in registers (DCAP), then a fixed-order warp + cross-warp reduction.
// gradx_tile_kernel (more rows) follows below.
// ---------------------------------------------------------------------------------------------------------------------
// The warps of a row take interleaved 32-point slices; their sums are added in warp order (fixed).
constexpr int GX_WPC = 4;

template <int KIND, int DCAP, int TWO>
__global__ void __launch_bounds__(256) gradx_kernel(const double *__restrict__ XcT, int ldc, int n_c, const double *__restrict__ XT, int ldx,
                                                    int n, int d, double variance, const double *__restrict__ inv_ls,
                                                    const double *__restrict__ G1, int ldg1, double s1, int add_t,
                                                    const double *__restrict__ G2, int ldg2, double s2,
                                                    double *__restrict__ out1, double *__restrict__ out2, int ldo) {
  constexpr int NV = TWO ? 2 * DCAP : DCAP;
  __shared__ double part[8][NV];
  const int lane = threadIdx.x & 31;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int c = blockIdx.x * (8 / GX_WPC) + warp / GX_WPC, slice = warp % GX_WPC;
  const bool live = c < n_c;
  double xc[DCAP], a1[DCAP], a2[TWO ? DCAP : 1];
#pragma unroll
  for (int q = 0; q < DCAP; ++q) {
    xc[q] = (live && q < d) ? XcT[(size_t)q * ldc + c] : 0.0;
    a1[q] = 0.0;
    if (TWO) a2[q] = 0.0;
  }
  if (live) {
    for (int j = slice * 32 + lane; j < n; j += 32 * GX_WPC) {
      double df[DCAP];
      double r2 = 0.0;
#pragma unroll
      for (int q = 0; q < DCAP; ++q) {
        df[q] = (q < d) ? xc[q] - XT[(size_t)q * ldx + j] : 0.0;
        r2 = fma(df[q], df[q], r2);
      }
      double k, dk;
      cov_k_dk<KIND>(r2, variance, k, dk);
      double g1 = G1[(size_t)c * ldg1 + j];
      if (add_t) g1 += G1[(size_t)j * ldg1 + c];
      const double w1 = dk * g1;
      double w2 = 0.0;
      if (TWO) w2 = dk * G2[(size_t)c * ldg2 + j];
#pragma unroll
      for (int q = 0; q < DCAP; ++q) {
        a1[q] = fma(w1, df[q], a1[q]);
        if (TWO) a2[q] = fma(w2, df[q], a2[q]);
      }
    }
  }
#pragma unroll
  for (int q = 0; q < DCAP; ++q) {
    const double v1 = warp_sum(a1[q]);
    if (lane == 0) part[warp][q] = v1;
    if (TWO) {
      const double v2 = warp_sum(a2[q]);
      if (lane == 0) part[warp][DCAP + q] = v2;
    }
  }
  __syncthreads();
  // one thread per (row of this CTA, value): the GX_WPC slices in order
  for (int e = threadIdx.x; e < (8 / GX_WPC) * NV; e += 256) {
    const int rloc = e / NV, i = e - rloc * NV;
    const int cc = blockIdx.x * (8 / GX_WPC) + rloc;
    const int q = i < DCAP ? i : i - DCAP;
    if (cc < n_c && q < d) {
      double v = 0.0;
#pragma unroll
      for (int w = 0; w < GX_WPC; ++w) v += part[rloc * GX_WPC + w][i];
      if (i < DCAP) out1[(size_t)cc * ldo + q] = s1 * inv_ls[q] * v;
      else out2[(size_t)cc * ldo + q] = s2 * inv_ls[q] * v;
    }
  }
}

// More than 8 rows (n_c >= GX_TILE_MIN_ROWS) take the tiled version: the one-warp-per-row kernel above spends its time waiting for
// global loads (ncu r2w: long-scoreboard 6.2 cycles per issue, 8 warps per SM, FP64 pipe 15% busy; 2.8 ms for 2048 x 16384 pairs).
// Here a CTA of 8 warps owns 8 rows and walks chunks of 128 training points that a 3-stage cp.async ring brings into shared
// memory once for all 8 rows (coordinates) plus each row's weights; a cluster of GX_SLICES CTAs shares the 8 rows, every CTA of
// it taking every GX_SLICES-th chunk, and rank 0 adds the slices' sums in rank order through distributed shared memory -- no
// scratch buffer, no atomics, the same bits on every run.
constexpr int GX_SLICES = 4;
constexpr int GX_CH = 128;
constexpr int GX_ROWS = 8;
constexpr int GX_TILE_MIN_ROWS = 9;       // every block of more than one 8-row group: a row's bits do not depend on the block it came in

template <int DCAP, int TWO>
struct GxTile {
  static constexpr int ST = DCAP > 16 ? 2 : 3;
  static constexpr int NV = TWO ? 2 * DCAP : DCAP;
  static constexpr int STAGE = DCAP * GX_CH + GX_ROWS * GX_CH * (TWO ? 2 : 1);      // doubles
  static constexpr size_t SMEM = (size_t)(ST * STAGE + GX_ROWS * NV) * sizeof(double);
};

template <int KIND, int DCAP, int TWO, int UNR>
__global__ void __launch_bounds__(256) gradx_tile_kernel(const double *__restrict__ XcT, int ldc, int n_c, const double *__restrict__ XT, int ldx,
                                                                      int n, int d, double variance, const double *__restrict__ inv_ls,
                                                                      const double *__restrict__ G1, int ldg1, double s1, int add_t,
                                                                      const double *__restrict__ G2, int ldg2, double s2,
                                                                      double *__restrict__ out1, double *__restrict__ out2, int ldo) {
  using T = GxTile<DCAP, TWO>;
  constexpr int NV = T::NV, ST = T::ST;
  extern __shared__ double gx_sm[];
  double *part = gx_sm + ST * T::STAGE;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int group = blockIdx.x / GX_SLICES;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int c = group * GX_ROWS + warp;
  const bool live = c < n_c;
  const int n_chunks = (n + GX_CH - 1) / GX_CH;
  const int mine = rank < n_chunks ? (n_chunks - rank + GX_SLICES - 1) / GX_SLICES : 0;

  auto issue = [&](int i) {
    if (i < mine) {
      double *sx = gx_sm + (i % ST) * T::STAGE, *sg1 = sx + DCAP * GX_CH, *sg2 = sg1 + GX_ROWS * GX_CH;
      const int j0 = (rank + GX_SLICES * i) * GX_CH;
      for (int e = tid; e < DCAP * GX_CH; e += 256) {
        const int q = e / GX_CH, j = j0 + (e % GX_CH);
        if (q < d && j < n) __pipeline_memcpy_async(sx + e, XT + (size_t)q * ldx + j, 8);
        else sx[e] = 0.0;
      }
      for (int e = tid; e < GX_ROWS * GX_CH; e += 256) {
        const int cc = group * GX_ROWS + e / GX_CH, j = j0 + (e % GX_CH);
        const bool in = cc < n_c && j < n;                    // weight 0 outside: the pair adds nothing
        if (in && !add_t) __pipeline_memcpy_async(sg1 + e, G1 + (size_t)cc * ldg1 + j, 8);
        else sg1[e] = in ? G1[(size_t)cc * ldg1 + j] + G1[(size_t)j * ldg1 + cc] : 0.0;
        if (TWO) {
          if (in) __pipeline_memcpy_async(sg2 + e, G2 + (size_t)cc * ldg2 + j, 8);
          else sg2[e] = 0.0;
        }
      }
    }
    __pipeline_commit();
  };

  double xc[DCAP], a1[DCAP], a2[TWO ? DCAP : 1];
#pragma unroll
  for (int q = 0; q < DCAP; ++q) {
    xc[q] = (live && q < d) ? XcT[(size_t)q * ldc + c] : 0.0;
    a1[q] = 0.0;
    if (TWO) a2[q] = 0.0;
  }
  for (int i = 0; i < ST - 1; ++i) issue(i);
  for (int i = 0; i < mine; ++i) {
    issue(i + ST - 1);
    __pipeline_wait_prior(ST - 1);
    __syncthreads();
    if (live) {
      const double *sx = gx_sm + (i % ST) * T::STAGE, *sg1 = sx + DCAP * GX_CH + warp * GX_CH, *sg2 = sg1 + GX_ROWS * GX_CH;
#pragma unroll UNR
      for (int u = 0; u < GX_CH / 32; ++u) {
        const int jj = lane + 32 * u;
        double df[DCAP];
        double ra = 0.0, rb = 0.0;                      // two chains: half the dependent latency
#pragma unroll
        for (int q = 0; q < DCAP; q += 2) {
          const double da = xc[q] - sx[q * GX_CH + jj];
          const double db = (q + 1 < DCAP) ? xc[q + 1] - sx[(q + 1) * GX_CH + jj] : 0.0;
          df[q] = da;
          if (q + 1 < DCAP) df[q + 1] = db;
          ra = fma(da, da, ra);
          rb = fma(db, db, rb);
        }
        double k, dk;
        cov_k_dk<KIND>(ra + rb, variance, k, dk);
        const double w1 = dk * sg1[jj];
        const double w2 = TWO ? dk * sg2[jj] : 0.0;
#pragma unroll
        for (int q = 0; q < DCAP; ++q) {
          a1[q] = fma(w1, df[q], a1[q]);
          if (TWO) a2[q] = fma(w2, df[q], a2[q]);
        }
      }
    }
    __syncthreads();
  }
  __pipeline_wait_prior(0);
#pragma unroll
  for (int q = 0; q < DCAP; ++q) {
    const double v1 = warp_sum(a1[q]);
    if (lane == 0) part[warp * NV + q] = v1;
    if (TWO) {
      const double v2 = warp_sum(a2[q]);
      if (lane == 0) part[warp * NV + DCAP + q] = v2;
    }
  }
  cluster.sync();
  if (rank == 0) {
    for (int e = tid; e < GX_ROWS * NV; e += 256) {
      const int w = e / NV, i = e - w * NV;
      const int cc = group * GX_ROWS + w;
      const int q = i < DCAP ? i : i - DCAP;
      if (cc < n_c && q < d) {
        double v = 0.0;
#pragma unroll
        for (int r = 0; r < GX_SLICES; ++r) v += cluster.map_shared_rank(part, r)[e];
        if (i < DCAP) out1[(size_t)cc * ldo + q] = s1 * inv_ls[q] * v;
        else out2[(size_t)cc * ldo + q] = s2 * inv_ls[q] * v;
      }
    }
  }
  cluster.sync();      // the other ranks' shared memory must outlive rank 0's reads
}

template <int KIND, int DCAP, int TWO, int UNR>
static int launch_gradx_tile(const double *XcT, int ldc, int n_c, const double *XT, int ldx, int n, int d, double variance,
                             const double *inv_ls, const double *G1, int ldg1, double s1, int add_t, const double *G2, int ldg2,
                             double s2, double *out1, double *out2, int ldo, cudaStream_t s) {
  using T = GxTile<DCAP, TWO>;
  auto kernel = gradx_tile_kernel<KIND, DCAP, TWO, UNR>;
  static FuncConfigMask configured;
  {
    FuncConfigOnce once(configured);
    if (once.needed) GPB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T::SMEM));
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(GX_SLICES * ((n_c + GX_ROWS - 1) / GX_ROWS)));
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = T::SMEM;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = GX_SLICES;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  GPB_CUDA(cudaLaunchKernelEx(&cfg, kernel, XcT, ldc, n_c, XT, ldx, n, d, variance, inv_ls, G1, ldg1, s1, add_t, G2, ldg2, s2, out1, out2, ldo));
  count_launch();
  return 0;
}

// 32 < d <= 64: the register version above would need 4 x 64 doubles per lane (the DCAP = 64 instantiation spilled 9 KB per thread).
// Here the candidate's coordinates sit in shared memory (read by broadcast), the squared distance is formed over all dimensions
// first, and the gradient sums are accumulated for 32 dimensions at a time in two sweeps over the training points.
template <int KIND, int TWO>
__global__ void __launch_bounds__(256) gradx_wide_kernel(const double *__restrict__ XcT, int ldc, int n_c, const double *__restrict__ XT,
                                                         int ldx, int n, int d, double variance, const double *__restrict__ inv_ls,
                                                         const double *__restrict__ G1, int ldg1, double s1, int add_t,
                                                         const double *__restrict__ G2, int ldg2, double s2,
                                                         double *__restrict__ out1, double *__restrict__ out2, int ldo) {
  __shared__ double xcs[8][64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * 8 + warp;
  if (c < n_c)
    for (int q = lane; q < 64; q += 32) xcs[warp][q] = (q < d) ? XcT[(size_t)q * ldc + c] : 0.0;
  __syncwarp();
  if (c >= n_c) return;
  const double *xc = xcs[warp];
  for (int q0 = 0; q0 < d; q0 += 32) {
    double a1[32], a2[TWO ? 32 : 1];
#pragma unroll
    for (int q = 0; q < 32; ++q) {
      a1[q] = 0.0;
      if (TWO) a2[q] = 0.0;
    }
    for (int j = lane; j < n; j += 32) {
      double r2 = 0.0;
      for (int q = 0; q < d; ++q) {
        const double df = xc[q] - XT[(size_t)q * ldx + j];
        r2 = fma(df, df, r2);
      }
      double k, dk;
      cov_k_dk<KIND>(r2, variance, k, dk);
      double g1 = G1[(size_t)c * ldg1 + j];
      if (add_t) g1 += G1[(size_t)j * ldg1 + c];
      const double w1 = dk * g1;
      double w2 = 0.0;
      if (TWO) w2 = dk * G2[(size_t)c * ldg2 + j];
#pragma unroll
      for (int q = 0; q < 32; ++q) {
        const double df = (q0 + q < d) ? xc[q0 + q] - XT[(size_t)(q0 + q) * ldx + j] : 0.0;
        a1[q] = fma(w1, df, a1[q]);
        if (TWO) a2[q] = fma(w2, df, a2[q]);
      }
    }
#pragma unroll
    for (int q = 0; q < 32; ++q) {
      if (q0 + q < d) {
        const double v1 = warp_sum(a1[q]);
        if (lane == 0) out1[(size_t)c * ldo + q0 + q] = s1 * inv_ls[q0 + q] * v1;
        if (TWO) {
          const double v2 = warp_sum(a2[q]);
          if (lane == 0) out2[(size_t)c * ldo + q0 + q] = s2 * inv_ls[q0 + q] * v2;
        }
      }
    }
  }
}

template <int KIND>
static int launch_gradx_wide(const double *XcT, int ldc, int n_c, const double *XT, int ldx, int n, int d, double variance,
                             const double *inv_ls, const double *G1, int ldg1, double s1, int add_t, const double *G2, int ldg2,
                             double s2, double *out1, double *out2, int ldo, cudaStream_t s) {
  const int blocks = (n_c + 7) / 8;
  if (G2)
    gradx_wide_kernel<KIND, 1><<<blocks, 256, 0, s>>>(XcT, ldc, n_c, XT, ldx, n, d, variance, inv_ls, G1, ldg1, s1, add_t, G2, ldg2, s2, out1, out2, ldo);
  else
    gradx_wide_kernel<KIND, 0><<<blocks, 256, 0, s>>>(XcT, ldc, n_c, XT, ldx, n, d, variance, inv_ls, G1, ldg1, s1, add_t, nullptr, 0, 0.0, out1, nullptr, ldo);
  count_launch();
  GPB_CHECK_LAUNCH();
  return 0;
}

template <int KIND, int DCAP>
static int launch_gradx_t(const double *XcT, int ldc, int n_c, const double *XT, int ldx, int n, int d, double variance,
                          const double *inv_ls, const double *G1, int ldg1, double s1, int add_t, const double *G2, int ldg2,
                          double s2, double *out1, double *out2, int ldo, cudaStream_t s) {
  const int blocks = (n_c + 8 / GX_WPC - 1) / (8 / GX_WPC);
  static const bool tile_on = !getenv("GPB_GRADX_TILE") || atoi(getenv("GPB_GRADX_TILE")) != 0;
  if (tile_on && n_c >= GX_TILE_MIN_ROWS) {
    constexpr int UNR = DCAP <= 16 ? 4 : 1;      // points of a chunk in flight per lane (0.72 -> 0.64 ms per block at D = 16; registers allow it up to 16)
    if (G2) return launch_gradx_tile<KIND, DCAP, 1, UNR>(XcT, ldc, n_c, XT, ldx, n, d, variance, inv_ls, G1, ldg1, s1, add_t, G2, ldg2, s2, out1, out2, ldo, s);
    return launch_gradx_tile<KIND, DCAP, 0, UNR>(XcT, ldc, n_c, XT, ldx, n, d, variance, inv_ls, G1, ldg1, s1, add_t, nullptr, 0, 0.0, out1, nullptr, ldo, s);
  }
  if (G2)
    gradx_kernel<KIND, DCAP, 1><<<blocks, 256, 0, s>>>(XcT, ldc, n_c, XT, ldx, n, d, variance, inv_ls, G1, ldg1, s1, add_t, G2, ldg2, s2, out1, out2, ldo);
  else
    gradx_kernel<KIND, DCAP, 0><<<blocks, 256, 0, s>>>(XcT, ldc, n_c, XT, ldx, n, d, variance, inv_ls, G1, ldg1, s1, add_t, nullptr, 0, 0.0, out1, nullptr, ldo);
  count_launch();
  GPB_CHECK_LAUNCH();
  return 0;
}

int launch_gradx(int kind, const double *XcT, int ldc, int n_c, const double *XT, int ldx, int n, int d, double variance,
                 const double *inv_ls, const double *G1, int ldg1, double s1, int add_t, const double *G2, int ldg2, double s2,
                 double *out1, double *out2, int ldo, cudaStream_t s) {
  if (n_c == 0) return 0;
  GPB_REQUIRE(d <= 64, "gradients_X: input dimension %d > 64 not supported", d);
#define GPB_GX(K_, D_) return launch_gradx_t<K_, D_>(XcT, ldc, n_c, XT, ldx, n, d, variance, inv_ls, G1, ldg1, s1, add_t, G2, ldg2, s2, out1, out2, ldo, s)
  if (kind == GPB_KERN_RBF) {
    if (d <= 4) GPB_GX(GPB_KERN_RBF, 4);
    if (d <= 8) GPB_GX(GPB_KERN_RBF, 8);
    if (d <= 16) GPB_GX(GPB_KERN_RBF, 16);
    if (d <= 32) GPB_GX(GPB_KERN_RBF, 32);
    return launch_gradx_wide<GPB_KERN_RBF>(XcT, ldc, n_c, XT, ldx, n, d, variance, inv_ls, G1, ldg1, s1, add_t, G2, ldg2, s2, out1, out2, ldo, s);
  } else {
    if (d <= 4) GPB_GX(GPB_KERN_MATERN52, 4);
    if (d <= 8) GPB_GX(GPB_KERN_MATERN52, 8);
    if (d <= 16) GPB_GX(GPB_KERN_MATERN52, 16);
    if (d <= 32) GPB_GX(GPB_KERN_MATERN52, 32);
    return launch_gradx_wide<GPB_KERN_MATERN52>(XcT, ldc, n_c, XT, ldx, n, d, variance, inv_ls, G1, ldg1, s1, add_t, G2, ldg2, s2, out1, out2, ldo, s);
  }
#undef GPB_GX
}

}  // namespace gpb
