// Blocked fp64 Cholesky / triangular inverse / SPD inverse / solves for ExactGaussianInference (north_star (b)).
//
// Replaces, on the device, GPy/GPy/util/linalg.py: jitchol's dpotrf (:56-60), pdinv (:193-214: dpotrf + dtrtri + dpotri +
// symmetrify), dpotrs (:116-125) and the logdet (:208).
//
// Design (B200 first, not a LAPACK transliteration): a recursive 2x2 splitting that produces L and L^-1 together,
//     [A11      ]      L11, M11 = cholinv(A11)
//     [A21  A22 ]      L21 = A21 M11^T          (GEMM, triangular k-range)   -- TRSM replaced by a product with L11^-1
//                      A22 -= L21 L21^T         (SYRK as lower-tile GEMM)
//                      L22, M22 = cholinv(A22)
//                      M21 = -M22 (L21 M11)     (two GEMMs, triangular k-ranges)
// so that every flop above the 128x128 leaves runs in the DMMA GEMM engine (gpb_gemm.cu) on large, regular tiles, and
// Ky^-1 = M^T M is one more lower-tile GEMM.  Flops: N^3/3 (L) + N^3/3 (M) + N^3/3 (M^T M) = N^3, the same count as
// dpotrf + dpotri.  The leaf factors a 128x128 block and inverts its factor in ONE fused rank-1 sweep held in registers.
#include "gpb_common.cuh"

namespace gpb {

// ---------------------------------------------------------------------------------------------------------------------
// Leaf: L = chol(A_blk), W = L^-1 for one 128x128 diagonal block, fused.
//
// Outer-product Cholesky applied to the augmented matrix [A | I]: the row operations that reduce A to L^T turn I into L^-1.
// The strictly-upper triangle of the working array holds W^T (it is free, A is symmetric), so step j is ONE rank-1 update
//     s[x][y] -= w[x] * u[y] / pivot_j     for y > j and (x <= j or y <= x),
// with u = column j of s, w = u with the diagonal entry replaced by 1.  Scaling by 1/sqrt(pivot) is deferred to the end.
// Each of the 256 threads keeps its 64 entries (x = ty + 16a, y = tx + 16b) in registers for all 128 steps; the only
// shared-memory traffic per step is the broadcast of column j (double buffered -> one __syncthreads per step).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int LEAF_THREADS = 256;
constexpr int LEAF_LD = TILE + 1;

// The 16 steps j = 16 JB + jt of one column block.  JB is a template parameter so that for every register entry (a, b)
// the membership test  y > j && (x <= j || y <= x)  collapses at compile time to at most three per-thread booleans
// (tx > jt, ty <= jt, tx <= ty): the body is straight-line predicated DFMAs, no per-entry index arithmetic or branches.
template <int MODE, int JB>
__device__ __forceinline__ void leaf_sweep_block(double (&s)[8][8], double *colbuf, double *piv, int tx, int ty, int tid,
                                                 int index_base, int *info, bool &failed) {
  const bool t_le = tx <= ty;  // y <= x inside a diagonal (a == b) register block
#pragma unroll 1
  for (int jt = 0; jt < 16; ++jt) {
    const int j = JB * 16 + jt;
    double *cb = colbuf + (j & 1) * TILE;
    if (tx == jt) {
#pragma unroll
      for (int a = 0; a < 8; ++a) cb[ty + 16 * a] = s[a][JB];
    }
    __syncthreads();
    const double pivot = cb[j];
    if (!(pivot > 0.0) && !failed) {
      failed = true;
      if (tid == 0) atomicCAS(info, 0, index_base + j + 1);
    }
    // the epilogue divides by sqrt(piv): Cholesky pivots are L_jj^2, in MODE 1 the diagonal entry is L_jj itself
    if (tid == 0) piv[j] = MODE ? pivot * pivot : pivot;
    const double rinv = 1.0 / pivot;
    const bool y_gt = tx > jt;   // y > j inside column block JB
    const bool x_le = ty <= jt;  // x <= j inside row block JB
    double w[8], u[8];
#pragma unroll
    for (int a = 0; a < 8; ++a) {
      if (MODE && a > JB) continue;  // trtri-only: rows below the current block are never touched
      double c = cb[ty + 16 * a];
      if (a == JB && ty == jt) c = 1.0;
      w[a] = c * rinv;
    }
#pragma unroll
    for (int b = JB; b < 8; ++b) u[b] = cb[tx + 16 * b];
#pragma unroll
    for (int b = JB; b < 8; ++b) {
#pragma unroll
      for (int a = 0; a < 8; ++a) {
        bool act;
        if (a < JB) {
          act = true;                        // W^T rows: x <= j always
        } else if (a == JB) {
          if (MODE) act = x_le;
          else act = x_le || (b == a && t_le);   // x > j: trailing matrix, needs y <= x (impossible for b > a)
        } else {
          if (MODE || b > a) continue;       // x > j: only the lower part of the trailing matrix (Cholesky mode)
          act = (b < a) || t_le;
        }
        if (b == JB) act = act && y_gt;
        if (act) s[a][b] = fma(-w[a], u[b], s[a][b]);
      }
    }
  }
}

// MODE 0: Cholesky + inverse (A is overwritten by L).  MODE 1: A already holds a lower-triangular factor L; only W = L^-1
// is produced (forward elimination on [L | I]: the same sweep restricted to the W^T rows, pivot = L_jj).
template <int MODE>
__global__ void __launch_bounds__(LEAF_THREADS, 1)
leaf_potrf_inv_kernel(double *__restrict__ A, int lda, double *__restrict__ Mi, int ldm, int index_base, int *info) {
  extern __shared__ double sm[];
  double *colbuf = sm;               // 2 x 128
  double *piv = sm + 2 * TILE;       // 128
  double *stage = sm + 3 * TILE;     // 128 x 129 (epilogue only)

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;

  double s[8][8];
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    const int x = ty + 16 * a;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      const int y = tx + 16 * b;
      s[a][b] = (y <= x) ? A[(size_t)x * lda + y] : 0.0;
    }
  }

  bool failed = false;
  leaf_sweep_block<MODE, 0>(s, colbuf, piv, tx, ty, tid, index_base, info, failed);
  leaf_sweep_block<MODE, 1>(s, colbuf, piv, tx, ty, tid, index_base, info, failed);
  leaf_sweep_block<MODE, 2>(s, colbuf, piv, tx, ty, tid, index_base, info, failed);
  leaf_sweep_block<MODE, 3>(s, colbuf, piv, tx, ty, tid, index_base, info, failed);
  leaf_sweep_block<MODE, 4>(s, colbuf, piv, tx, ty, tid, index_base, info, failed);
  leaf_sweep_block<MODE, 5>(s, colbuf, piv, tx, ty, tid, index_base, info, failed);
  leaf_sweep_block<MODE, 6>(s, colbuf, piv, tx, ty, tid, index_base, info, failed);
  leaf_sweep_block<MODE, 7>(s, colbuf, piv, tx, ty, tid, index_base, info, failed);
  __syncthreads();

  // ---- epilogue: scale, stage through shared memory, coalesced writes of L (into A) and W = L^-1 (into Mi) ----
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) stage[(ty + 16 * a) * LEAF_LD + tx + 16 * b] = s[a][b];
  __syncthreads();
  for (int e = tid; e < TILE * TILE; e += LEAF_THREADS) {
    const int r = e >> 7, c = e & 127;
    double l, wv;
    if (c < r) {
      l = stage[r * LEAF_LD + c] / sqrt(piv[c]);
      wv = stage[c * LEAF_LD + r] / sqrt(piv[r]);
    } else if (c == r) {
      const double d = sqrt(piv[r]);
      l = d;
      wv = 1.0 / d;
    } else {
      l = 0.0;
      wv = 0.0;
    }
    if (!MODE) A[(size_t)r * lda + c] = l;
    Mi[(size_t)r * ldm + c] = wv;
  }
}

static int launch_leaf(Factor &f, int off, int mode) {
  const size_t smem = (size_t)(3 * TILE + TILE * LEAF_LD) * sizeof(double);
  static bool configured = false;
  if (!configured) {
    GPB_CUDA(cudaFuncSetAttribute(leaf_potrf_inv_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GPB_CUDA(cudaFuncSetAttribute(leaf_potrf_inv_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  double *a = f.A + (size_t)off * f.np + off;
  double *m = f.Mi + (size_t)off * f.np + off;
  if (mode == 0)
    leaf_potrf_inv_kernel<0><<<1, LEAF_THREADS, smem, f.stream>>>(a, f.np, m, f.np, off, f.info);
  else
    leaf_potrf_inv_kernel<1><<<1, LEAF_THREADS, smem, f.stream>>>(a, f.np, m, f.np, off, f.info);
  count_launch();
  GPB_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------------------------------
__global__ void copy2d_kernel(double *__restrict__ dst, int ldd, const double *__restrict__ src, int lds, int rows,
                              int cols2) {
  // cols2 = cols / 2 (16-byte elements)
  const size_t total = (size_t)rows * cols2;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(e / cols2), c = (int)(e - (size_t)r * cols2);
    reinterpret_cast<double2 *>(dst + (size_t)r * ldd)[c] = reinterpret_cast<const double2 *>(src + (size_t)r * lds)[c];
  }
}

int launch_copy2d(double *dst, int ldd, const double *src, int lds, int rows, int cols, cudaStream_t s) {
  if (rows == 0 || cols == 0) return 0;
  GPB_REQUIRE(cols % 2 == 0 && ldd % 2 == 0 && lds % 2 == 0, "copy2d: even sizes only");
  const size_t total = (size_t)rows * (cols / 2);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  copy2d_kernel<<<blocks, 256, 0, s>>>(dst, ldd, src, lds, rows, cols / 2);
  count_launch();
  GPB_CHECK_LAUNCH();
  return 0;
}

// A[i][j] = A[j][i] for j > i (GPy symmetrify, linalg_cython.pyx:9-18), tiled through shared memory.
__global__ void symmetrize_lower_kernel(double *A, int ld, int n) {
  __shared__ double t[32][33];
  const int bi = blockIdx.y, bj = blockIdx.x;  // source tile (bi, bj) with bj <= bi, written to (bj, bi)
  if (bj > bi) return;
  const int tx = threadIdx.x, ty = threadIdx.y;
  for (int r = ty; r < 32; r += 8) {
    const int i = bi * 32 + r, j = bj * 32 + tx;
    t[r][tx] = (i < n && j < n) ? A[(size_t)i * ld + j] : 0.0;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int i = bj * 32 + r, j = bi * 32 + tx;  // destination (i, j) = source (j, i)
    if (i < n && j < n && j > i) A[(size_t)i * ld + j] = t[tx][r];
  }
}

int launch_symmetrize_lower(double *A, int ld, int n, cudaStream_t s) {
  const int nb = (n + 31) / 32;
  symmetrize_lower_kernel<<<dim3(nb, nb), dim3(32, 8), 0, s>>>(A, ld, n);
  count_launch();
  GPB_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// potrf + triangular inverse (recursive)
// ---------------------------------------------------------------------------------------------------------------------
static int cholinv(Factor &f, int off, int n) {
  if (n == TILE) return launch_leaf(f, off, 0);
  const int ld = f.np;
  const int h = ((n / TILE) / 2) * TILE;  // first half (multiple of 128), second half n - h >= h
  const int r = n - h;
  GPB_TRY(cholinv(f, off, h));
  double *A21 = f.A + (size_t)(off + h) * ld + off;
  double *A22 = f.A + (size_t)(off + h) * ld + off + h;
  double *M11 = f.Mi + (size_t)off * ld + off;
  double *M21 = f.Mi + (size_t)(off + h) * ld + off;
  double *M22 = f.Mi + (size_t)(off + h) * ld + off + h;
  double *S21 = f.W + (size_t)(off + h) * ld + off;  // scratch with the shape of the (2,1) block
  GemmArgs g;
  // S21 = A21 * M11^T      (M11 lower: k <= column tile)
  g = GemmArgs{A21, ld, M11, ld, S21, ld, r, h, h, 1.0, 0.0, 0, 0, 2};
  GPB_TRY(gemm_launch(LAYOUT_ROWK, LAYOUT_ROWK, g, f.stream));
  GPB_TRY(launch_copy2d(A21, ld, S21, ld, r, h, f.stream));  // A21 <- L21
  // A22 -= L21 * L21^T     (lower tiles)
  g = GemmArgs{S21, ld, S21, ld, A22, ld, r, r, h, -1.0, 1.0, 1, 0, 0};
  GPB_TRY(gemm_launch(LAYOUT_ROWK, LAYOUT_ROWK, g, f.stream));
  GPB_TRY(cholinv(f, off + h, r));
  // S21 = L21 * M11        (M11 lower: k >= column tile)
  g = GemmArgs{A21, ld, M11, ld, S21, ld, r, h, h, 1.0, 0.0, 0, 2, 0};
  GPB_TRY(gemm_launch(LAYOUT_ROWK, LAYOUT_COLK, g, f.stream));
  // M21 = -M22 * S21       (M22 lower: k <= row tile)
  g = GemmArgs{M22, ld, S21, ld, M21, ld, r, h, r, -1.0, 0.0, 0, 0, 1};
  GPB_TRY(gemm_launch(LAYOUT_ROWK, LAYOUT_COLK, g, f.stream));
  return 0;
}

// M = L^-1 for a lower-triangular factor already stored in f.A (dtrtri, linalg.py:217-227): the inverse half of cholinv.
static int trtri_rec(Factor &f, int off, int n) {
  if (n == TILE) return launch_leaf(f, off, 1);
  const int ld = f.np;
  const int h = ((n / TILE) / 2) * TILE;
  const int r = n - h;
  GPB_TRY(trtri_rec(f, off, h));
  GPB_TRY(trtri_rec(f, off + h, r));
  double *A21 = f.A + (size_t)(off + h) * ld + off;
  double *M11 = f.Mi + (size_t)off * ld + off;
  double *M21 = f.Mi + (size_t)(off + h) * ld + off;
  double *M22 = f.Mi + (size_t)(off + h) * ld + off + h;
  double *S21 = f.W + (size_t)(off + h) * ld + off;
  GemmArgs g{A21, ld, M11, ld, S21, ld, r, h, h, 1.0, 0.0, 0, 2, 0};
  GPB_TRY(gemm_launch(LAYOUT_ROWK, LAYOUT_COLK, g, f.stream));
  g = GemmArgs{M22, ld, S21, ld, M21, ld, r, h, r, -1.0, 0.0, 0, 0, 1};
  GPB_TRY(gemm_launch(LAYOUT_ROWK, LAYOUT_COLK, g, f.stream));
  return 0;
}

int factor_trtri(Factor &f) {
  GPB_REQUIRE(f.np % TILE == 0 && f.np >= TILE, "factor: padded size must be a multiple of 128");
  GPB_CUDA(cudaMemsetAsync(f.info, 0, sizeof(int), f.stream));
  return trtri_rec(f, 0, f.np);
}

int factor_potrf_inv(Factor &f) {
  GPB_REQUIRE(f.np % TILE == 0 && f.np >= TILE, "factor: padded size must be a multiple of 128");
  GPB_CUDA(cudaMemsetAsync(f.info, 0, sizeof(int), f.stream));
  return cholinv(f, 0, f.np);
}

// Ky^-1 = M^T M: W[i][j] = sum_{k >= max(i,j)} M[k][i] M[k][j]; lower tiles only (diagonal tiles complete).
int factor_potri(Factor &f) {
  GemmArgs g{f.Mi, f.np, f.Mi, f.np, f.W, f.np, f.np, f.np, f.np, 1.0, 0.0, 1, 1, 0};
  return gemm_launch(LAYOUT_COLK, LAYOUT_COLK, g, f.stream);
}

// ---------------------------------------------------------------------------------------------------------------------
// triangular matrix-vector products with M = L^-1 (dpotrs replacement): z = M y, alpha = M^T z
// ---------------------------------------------------------------------------------------------------------------------
// z[i] = sum_{k <= i} M[i][k] y[k]: one warp per row, coalesced along the row; fixed reduction order.
__global__ void trmv_lower_kernel(const double *__restrict__ M, int ld, int np, const double *__restrict__ y,
                                  double *__restrict__ z) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= np) return;
  const int i = warp;
  const int kend = (i / TILE + 1) * TILE;  // the diagonal leaf block holds explicit zeros above the diagonal
  const double *row = M + (size_t)i * ld;
  double acc = 0.0;
  for (int k = lane; k < kend; k += 32) acc = fma(row[k], y[k], acc);
  acc = warp_sum(acc);
  if (lane == 0) z[i] = acc;
}

// partial[rc][j] = sum_{i in row chunk rc} M[i][j] z[i] for column block cb <= rc; thread per column.
__global__ void trmv_lower_t_partial_kernel(const double *__restrict__ M, int ld, const double *__restrict__ z,
                                            double *__restrict__ part, int np) {
  const int cb = blockIdx.x, rc = blockIdx.y;
  if (rc < cb) return;
  const int j = cb * TILE + threadIdx.x;
  const double *p = M + (size_t)(rc * TILE) * ld + j;
  const double *zz = z + rc * TILE;
  double acc = 0.0;
#pragma unroll 8
  for (int i = 0; i < TILE; ++i) acc = fma(p[(size_t)i * ld], zz[i], acc);
  part[(size_t)rc * np + j] = acc;
}

__global__ void trmv_lower_t_reduce_kernel(const double *__restrict__ part, int np, double *__restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= np) return;
  const int nb = np / TILE;
  double acc = 0.0;
  for (int rc = j / TILE; rc < nb; ++rc) acc += part[(size_t)rc * np + j];
  out[j] = acc;
}

int factor_solve(Factor &f, const double *Y, int p, double *z, double *alpha) {
  const int np = f.np, nb = np / TILE;
  for (int c = 0; c < p; ++c) {
    const double *y = Y + (size_t)c * np;
    double *zc = z + (size_t)c * np, *ac = alpha + (size_t)c * np;
    trmv_lower_kernel<<<(np * 32 + 255) / 256, 256, 0, f.stream>>>(f.Mi, np, np, y, zc);
    GPB_CHECK_LAUNCH();
    trmv_lower_t_partial_kernel<<<dim3(nb, nb), TILE, 0, f.stream>>>(f.Mi, np, zc, f.part, np);
    GPB_CHECK_LAUNCH();
    trmv_lower_t_reduce_kernel<<<(np + 255) / 256, 256, 0, f.stream>>>(f.part, np, ac);
    GPB_CHECK_LAUNCH();
    count_launch(3);
  }
  return 0;
}

// 2 sum_{i<n} log L_ii, single block, fixed order.
__global__ void logdet_kernel(const double *__restrict__ L, int ld, int n, double *out) {
  __shared__ double scratch[32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += 1024) acc += log(L[(size_t)i * ld + i]);
  acc = block_sum<1024>(acc, scratch);
  if (threadIdx.x == 0) *out = 2.0 * acc;
}

int factor_logdet(Factor &f, double *out_dev) {
  logdet_kernel<<<1, 1024, 0, f.stream>>>(f.A, f.np, f.n, out_dev);
  count_launch();
  GPB_CHECK_LAUNCH();
  return 0;
}

}  // namespace gpb
