// Launchers of the covariance / gradient / predictive kernels (gpb_kernels.cu, gpb_predict.cu).
#pragma once
#include "gpb_common.cuh"

namespace gpb {

constexpr int KTILE = 64;  // covariance / gradient-reduction tile edge (kmat_kernel, kgrad_kernel)
// doubles of scratch launch_kgrad needs for its per-tile partials (padded sizes are multiples of 128)
inline size_t kgrad_part_doubles(int rows_pad, int cols_pad, int d, int fused) {
  const size_t tr = rows_pad / KTILE, tc = cols_pad / KTILE;
  return (fused ? tr * (tr + 1) / 2 : tr * tc) * (size_t)(d + 2);
}

// XsT[q][i] = X[i][q] / scale[q] (zero padded to ldx); tag is informational (0 lengthscales, 1 Gower ranges)
int launch_scale_transpose(const double *X, int n, int d, const double *ls_dev, double *XsT, int ldx, cudaStream_t s, int tag = 0);

// mode: 0 rect exact, 1 padded Ky (identity outside n x n, diag_add on the diagonal), 2 rect zero padded,
//       3 = 1 but only the 128-blocks on / below the diagonal (what the factorisation reads)
// gflag != NULL: the Gower product kernel (coordinates pre-divided by the ranges on continuous dimensions, gflag[q] != 0 marks
// a discrete dimension, `variance` is variance^d)
int launch_kmat(int kind, const double *XaT, int lda, const double *XbT, int ldb, int d, int n_rows, int n_cols,
                double variance, double diag_add, int mode, double *out, int ldo, int rows_pad, int cols_pad,
                cudaStream_t s, const double *gflag = nullptr, int row_start = 0,    // row_start: only rows >= it (multiple of 64)
                const double *theta = nullptr);   // theta != NULL: variance = theta[0], diag_add = theta[1] read on the device
// out_dev[0] = sum K_gower . G  (variance-gradient term under the Gower patch); part: tiles doubles
int launch_kvar_gower(int kind, int fused, const double *XaT, int lda, const double *XbT, int ldb, int d, int n_rows, int n_cols,
                      double vpow, const double *gflag, const double *G, int ldg, const double *alpha, int ld_alpha, int p_out,
                      double *part, double *out_dev, cudaStream_t s);

// out_dev[0] = sum K.G, out_dev[1] = tr G (fused only), out_dev[2+q] = sum (k'/r) G ds_q^2.
// part: scratch of tiles * (d + 2) doubles.
int launch_kgrad(int kind, int fused, const double *XaT, int lda, const double *XbT, int ldb, int d, int n_rows, int n_cols,
                 double variance, const double *G, int ldg, const double *alpha, int ld_alpha, int p_out, double *part,
                 double *out_dev, cudaStream_t s, const double *theta = nullptr);   // theta != NULL: variance = theta[0] on the device

int launch_gradx(int kind, const double *XcT, int ldc, int n_c, const double *XT, int ldx, int n, int d, double variance,
                 const double *inv_ls, const double *G1, int ldg1, double s1, int add_t, const double *G2, int ldg2, double s2,
                 double *out1, double *out2, int ldo, cudaStream_t s);

// gpb_predict.cu
// mu[c][p] = sum_n KxT[c][n] alpha_p[n]   (row dot products, one warp per candidate)
int launch_rowdot(const double *KxT, int ld, int n_c, int n, const double *alpha, int ld_alpha, int p, double *mu, cudaStream_t s);
// var[c] = base - sum_n Vt[c][n]^2
int launch_var_from_vt(const double *Vt, int ld, int n_c, int n, double base, double *var, cudaStream_t s);
// n_c <= 8 candidates: mu (KxT), var = var_base - sum Vt^2 (Vt), dmu (want_g) and dvar (want_g, Ut) with the training points split
// over many CTAs; part: skinny_moments_part_doubles(n_c, np, d) doubles of scratch
size_t skinny_moments_part_doubles(int n_c, int np, int d);
int launch_skinny_moments(int kind, const double *KxT, const double *Vt, const double *Ut, int ld, int n_c, int n, const double *alpha,
                          const double *XcT, int ldc, const double *XT, int ldx, int d, double variance, const double *inv_ls,
                          double var_base, int want_g, double *part, double *mu, double *var, double *dmu, double *dvar,
                          cudaStream_t s);
// gpb_skinny.cu: the same call (n_c <= 8, one output column, stationary kernel) as ONE persistent cooperative kernel -- covariance row,
// Z = M k*, U = M^T Z and all four reductions, two streaming passes over the triangle of M, 16-byte loads, grid-wide barriers
size_t skinny_fused_part12_doubles(int np);
size_t skinny_fused_part3_doubles(int d);
int launch_skinny_fused(int kind, const double *M, int np, int n, int d, int mc, int level, const double *XT, const double *Xc,
                        const double *ls, const double *inv_ls, const double *alpha, double variance, double var_base, double *Kx,
                        double *Dk, double *Z, double *U, double *part12, double *part3, double *mu, double *var, double *dmu,
                        double *dvar, cudaStream_t s);
// GPModel.predict clip + get_quantiles + EI/LCB (+ gradients) + AcquisitionBase sign
int launch_acq_epilogue(int acq, double par, double fmin, int n_c, int d, const double *mu, const double *var, const double *dmu,
                        const double *dvar, double *f, double *df, double *mean_out, double *sd_out, double *dmdx_out,
                        double *dsdx_out, cudaStream_t s);
// AcquisitionLP (LP.py:70-132): log transform + hammer-function penalisers around nb batch points, value and gradient
int launch_lp_epilogue(int n_c, int d, int nb, const double *Xc, const double *Xb, const double *r, const double *s, int transform,
                       const double *F, const double *dF, double *f_out, double *df_out, cudaStream_t st);
// running top-k of the k smallest f (ties -> lowest index); state on device: vals[k], idx[k]
int launch_topk_init(double *vals, long long *idx, int k, cudaStream_t s);
int launch_topk_update(const double *f, int n_c, long long index_base, double *vals, long long *idx, int k, cudaStream_t s);
// rows [value, global index as a double, d coordinates] of the k slots into out (device, k x (d + 2)); empty slots -> [NaN, -1, NaN..]
int launch_topk_pack(const double *vals, const long long *idx, const double *Xc_dev, int d, int k, long long index_offset, double *out,
                     cudaStream_t s);
int launch_min(const double *v, int n, double *out, cudaStream_t s);
// out[0] = max_i |Ky alpha - y|_i / (|Ky| |alpha| + |y|)_i over the n x n leading part of the full symmetric Ky
int launch_solve_residual(const double *Ky, int ld, int n, const double *alpha, const double *y, double *out, cudaStream_t s);

}  // namespace gpb
