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
#include <stdlib.h>

#include <cooperative_groups.h>

#include "gpb_common.cuh"
#include "gpb_gemm_tile.cuh"

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

// One column block b of the rank-1 update of step j = 16 JB + jt.  Called from fully unrolled loops, so `b` and JB are
// compile-time after unrolling.  The membership test of entry (a, b),
//     y > j && (x <= j || y <= x)        (x = ty + 16 a, y = tx + 16 b),
// is folded into the OPERANDS instead of predicating 64 DFMAs: the caller zeroes u for the columns left of j, and passes
// row factors that are zero where the row is inactive (nw: plain -w; nwd: -w masked by y <= x for the diagonal register
// blocks; nw_jd / nw_jo: row block JB on / off the diagonal register block).  A zero factor leaves the entry unchanged
// (s + (-0 * u) == s), so the body is straight-line unconditional DFMAs.
template <int MODE, int JB>
__device__ __forceinline__ void leaf_update_col(double (&s)[8][8], const double (&nw)[8], const double (&nwd)[8], double nw_jd,
                                                double nw_jo, double ub, int b) {
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    if (a < JB) {
      s[a][b] = fma(nw[a], ub, s[a][b]);                     // W^T rows: x <= j always
    } else if (a == JB) {
      s[a][b] = fma(b == JB ? nw_jd : nw_jo, ub, s[a][b]);   // rows of the current block
    } else {
      if (MODE || b > a) continue;                           // x > j: only the lower part of the trailing matrix
      s[a][b] = fma(b == a ? nwd[a] : nw[a], ub, s[a][b]);
    }
  }
}

// ---- split-phase CTA barrier (mbarrier): arrive right after publishing a column, wait just before reading the next one ----
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("{\n.reg .b64 t;\nmbarrier.arrive.shared::cta.b64 t, [%0];\n}" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
  asm volatile(
      "{\n.reg .pred p;\nLEAF_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra LEAF_DONE;\nbra LEAF_WAIT;\nLEAF_DONE:\n}" ::"r"(
          (unsigned)__cvta_generic_to_shared(bar)),
      "r"(parity)
      : "memory");
}
// Zero-instruction scheduling fence for one value: everything that consumes v is ordered after this point.
__device__ __forceinline__ void pin(double &v) { asm volatile("" : "+d"(v)); }

// The 16 steps j = 16 JB + jt of one column block, software pipelined with a split-phase barrier.  Phase j of the mbarrier
// completes when all 256 threads have passed the point "column j is published".  A step is
//     wait(phase j) -> load column j (pivot, row factors, column factors) -> reciprocal -> update the register column block
//     that holds column j + 1 -> its 16 owner threads publish it -> arrive(phase j + 1) -> the bulk of the rank-1 update,
// so that the publish -> barrier -> load -> reciprocal chain of the next column (~180 cycles, scripts/microbench) runs under
// the bulk DFMAs of this one instead of after them.  `pin` keeps the compiler from hoisting the bulk above the publish.
// Two column buffers suffice: a thread overwrites buffer (j + 1) & 1 only after phase j completed, i.e. after every thread
// has finished its loads of column j - 1 (all loads of a step precede its arrive).
template <int MODE, int JB>
__device__ __forceinline__ void leaf_sweep_block(double (&s)[8][8], double *colbuf, double *piv, uint64_t *bar, int tx, int ty,
                                                 int tid, int &fail_at) {
  const bool t_le = tx <= ty;  // y <= x inside a diagonal (a == b) register block
#pragma unroll 1
  for (int jt = 0; jt < 16; ++jt) {
    const int j = JB * 16 + jt;
    const double *cb = colbuf + (j & 1) * TILE;
    double *nb = colbuf + ((j + 1) & 1) * TILE;
    mbar_wait(bar, j & 1);
    const double pivot = cb[j];
    if (!(pivot > 0.0) && fail_at == 0) fail_at = j + 1;
    // the epilogue divides by sqrt(piv): Cholesky pivots are L_jj^2, in MODE 1 the diagonal entry is L_jj itself
    if (tid == 0) piv[j] = MODE ? pivot * pivot : pivot;
    const double nrinv = -1.0 / pivot;
    const bool y_gt = tx > jt;   // y > j inside column block JB
    const bool x_le = ty <= jt;  // x <= j inside row block JB
    double nw[8], nwd[8], u[8];
#pragma unroll
    for (int a = 0; a < 8; ++a) {
      if (MODE && a > JB) continue;  // trtri-only: rows below the current block are never touched
      double c = cb[ty + 16 * a];
      if (a == JB && ty == jt) c = 1.0;
      nw[a] = c * nrinv;
      nwd[a] = t_le ? nw[a] : 0.0;
    }
    const double nw_jo = x_le ? nw[JB] : 0.0;                           // row block JB, columns right of the diagonal block
    const double nw_jd = (MODE ? x_le : (x_le || t_le)) ? nw[JB] : 0.0;  // row block JB, diagonal register block
#pragma unroll
    for (int b = JB; b < 8; ++b) u[b] = cb[tx + 16 * b];
    if (jt < 15) {
      // column j + 1 lives in column block JB of the threads with tx == jt + 1
      leaf_update_col<MODE, JB>(s, nw, nwd, nw_jd, nw_jo, y_gt ? u[JB] : 0.0, JB);
      if (tx == jt + 1) {
#pragma unroll
        for (int a = 0; a < 8; ++a) nb[ty + 16 * a] = s[a][JB];
      }
      mbar_arrive(bar);
#pragma unroll
      for (int b = JB + 1; b < 8; ++b) pin(u[b]);
#pragma unroll
      for (int b = JB + 1; b < 8; ++b) leaf_update_col<MODE, JB>(s, nw, nwd, nw_jd, nw_jo, u[b], b);
    } else {
      // last column of the block: nothing of block JB is right of it; column j + 1 is the first column of block JB + 1,
      // owned by the threads with tx == 0
      constexpr int NB = JB < 7 ? JB + 1 : 7;
      if (JB < 7) {
        leaf_update_col<MODE, JB>(s, nw, nwd, nw_jd, nw_jo, u[NB], NB);
        if (tx == 0) {
#pragma unroll
          for (int a = 0; a < 8; ++a) nb[ty + 16 * a] = s[a][NB];
        }
      }
      mbar_arrive(bar);
#pragma unroll
      for (int b = JB + 2; b < 8; ++b) pin(u[b]);
#pragma unroll
      for (int b = JB + 2; b < 8; ++b) leaf_update_col<MODE, JB>(s, nw, nwd, nw_jd, nw_jo, u[b], b);
    }
  }
}

// MODE 0: Cholesky + inverse (A is overwritten by L).  MODE 1: A already holds a lower-triangular factor L; only W = L^-1
// is produced (forward elimination on [L | I]: the same sweep restricted to the W^T rows, pivot = L_jj).
template <int MODE>
__global__ void __launch_bounds__(LEAF_THREADS, 1)
leaf_potrf_inv_kernel(double *__restrict__ A, int lda, double *__restrict__ Mi, int ldm, int index_base, int *info) {
  extern __shared__ double sm[];
  uint64_t *bar = reinterpret_cast<uint64_t *>(sm);  // sm[0]; sm[1] pads to 16 bytes
  double *colbuf = sm + 2;           // 2 x 128
  double *piv = colbuf + 2 * TILE;   // 128 (after the sweep: 1 / sqrt(pivot))
  double *stage = piv + TILE;        // 128 x 129 (epilogue only)

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  if (tid == 0) mbar_init(bar, LEAF_THREADS);

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

  if (tx == 0) {
#pragma unroll
    for (int a = 0; a < 8; ++a) colbuf[ty + 16 * a] = s[a][0];
  }
  __syncthreads();   // mbarrier initialised, column 0 published
  mbar_arrive(bar);  // phase 0
  int fail_at = 0;
  leaf_sweep_block<MODE, 0>(s, colbuf, piv, bar, tx, ty, tid, fail_at);
  leaf_sweep_block<MODE, 1>(s, colbuf, piv, bar, tx, ty, tid, fail_at);
  leaf_sweep_block<MODE, 2>(s, colbuf, piv, bar, tx, ty, tid, fail_at);
  leaf_sweep_block<MODE, 3>(s, colbuf, piv, bar, tx, ty, tid, fail_at);
  leaf_sweep_block<MODE, 4>(s, colbuf, piv, bar, tx, ty, tid, fail_at);
  leaf_sweep_block<MODE, 5>(s, colbuf, piv, bar, tx, ty, tid, fail_at);
  leaf_sweep_block<MODE, 6>(s, colbuf, piv, bar, tx, ty, tid, fail_at);
  leaf_sweep_block<MODE, 7>(s, colbuf, piv, bar, tx, ty, tid, fail_at);
  if (tid == 0 && fail_at != 0) atomicCAS(info, 0, index_base + fail_at);  // every thread saw the same pivots

  // ---- epilogue: scale, stage through shared memory, coalesced writes of L (into A) and W = L^-1 (into Mi) ----
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) stage[(ty + 16 * a) * LEAF_LD + tx + 16 * b] = s[a][b];
  __syncthreads();
  if (tid < TILE) piv[tid] = 1.0 / sqrt(piv[tid]);   // one reciprocal square root per column instead of two per element
  __syncthreads();
  for (int e = tid; e < TILE * TILE; e += LEAF_THREADS) {
    const int r = e >> 7, c = e & 127;
    double l, wv;
    if (c < r) {
      l = stage[r * LEAF_LD + c] * piv[c];
      wv = stage[c * LEAF_LD + r] * piv[r];
    } else if (c == r) {
      const double d = piv[r];
      l = 1.0 / d;
      wv = d;
    } else {
      l = 0.0;
      wv = 0.0;
    }
    if (!MODE) A[(size_t)r * lda + c] = l;
    Mi[(size_t)r * ldm + c] = wv;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Blocked leaf (MODE 0 only): the 128x128 block lives in shared memory as 8 x 8 sub-blocks of 16 x 16 and every flop outside
// the 16 x 16 diagonal factorisations is a DMMA.8x8x4 issued by one of the 8 warps:
//     for k = 0..7:   warp 0:     L_kk, M_kk = chol + inverse of the 16 x 16 diagonal block, in registers ([D | I] elimination,
//                                 one column per lane, pivot column broadcast by shuffles)
//                     warps 1-7:  meanwhile, block row k-1 of M:  M_(k-1)j = -M_(k-1)(k-1) sum_m L_(k-1)m M_mj
//                     panel       L_ik = A_ik M_kk^T                      (i > k, one sub-block per warp)
//                     trailing    A_ij -= L_ik L_jk^T                     (i >= j > k, sub-blocks round robin over the warps)
// The lower triangle of the shared array holds A -> L, the strictly-upper sub-blocks hold M^T, the diagonal sub-blocks of M sit
// in their own array.  Row stride 132 == 4 (mod 16): the 8 x 4 DMMA operand fragments are bank-conflict free in both
// orientations ([row][k] and [k][row]).  The serial chain is 128 pivots x ~100 cycles instead of 128 block-wide barriers.
// ---------------------------------------------------------------------------------------------------------------------
#ifdef GPB_LEAF_TIMING
__device__ long long g_leaf_clk[64];
#define LEAF_CLK(i) do { if (threadIdx.x == 0) g_leaf_clk[i] = clock64(); } while (0)
#else
#define LEAF_CLK(i) do { } while (0)
#endif
constexpr int LB = 16;            // sub-block size
constexpr int LNB = TILE / LB;    // 8 sub-blocks per side
constexpr int LSD = TILE + 4;     // row stride of the 128 x 128 array
constexpr int LMD = LB + 4;       // row stride of a 16 x 16 sub-block held separately

__device__ __forceinline__ void leaf_dmma(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// c (16 x 16, as 2 x 2 accumulator fragments) += sa * A * B.  A_KM: A is stored [k][m] instead of [m][k]; B_KN: B is stored
// [k][n] instead of [n][k].
template <bool A_KM, bool B_KN>
__device__ __forceinline__ void blk_mma(double (&c)[2][2][2], const double *__restrict__ A, int lda, const double *__restrict__ B,
                                        int ldb, double sa, int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int kk = 0; kk < LB; kk += 4) {
    double a[2], b[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      a[i] = sa * (A_KM ? A[(kk + t) * lda + 8 * i + g] : A[(8 * i + g) * lda + kk + t]);
      b[i] = B_KN ? B[(kk + t) * ldb + 8 * i + g] : B[(8 * i + g) * ldb + kk + t];
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) leaf_dmma(c[i][j][0], c[i][j][1], a[i], b[j]);
  }
}
__device__ __forceinline__ void blk_zero(double (&c)[2][2][2]) {
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) c[i][j][0] = c[i][j][1] = 0.0;
}
__device__ __forceinline__ void blk_load(double (&c)[2][2][2], const double *S, int ld, int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const double2 v = *reinterpret_cast<const double2 *>(S + (8 * i + g) * ld + 8 * j + 2 * t);
      c[i][j][0] = v.x;
      c[i][j][1] = v.y;
    }
}
__device__ __forceinline__ void blk_store(const double (&c)[2][2][2], double *S, int ld, int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) *reinterpret_cast<double2 *>(S + (8 * i + g) * ld + 8 * j + 2 * t) = make_double2(c[i][j][0], c[i][j][1]);
}
// S[n][m] = c(m, n)
__device__ __forceinline__ void blk_store_t(const double (&c)[2][2][2], double *S, int ld, int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      S[(8 * j + 2 * t) * ld + 8 * i + g] = c[i][j][0];
      S[(8 * j + 2 * t + 1) * ld + 8 * i + g] = c[i][j][1];
    }
}

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// fp64 reciprocal / reciprocal square root off the MUFU seed (~20 bits) + Newton steps: a few dependent DFMAs instead of the
// division / sqrt slow paths; the pivot chain of the diagonal sub-blocks is the critical path of the leaf.  Not correctly
// rounded (<= 1 ulp); non-positive or NaN inputs give garbage, which the caller has already flagged as "not positive definite".
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
}
__device__ __forceinline__ double fast_rsqrt(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#pragma unroll
  for (int it = 0; it < 3; ++it) {
    const double e = fma(-x * y, y, 1.0);
    y = fma(0.5 * y, e, y);
  }
  return y;
}

// One warp: D (16 x 16 symmetric positive definite, lower triangle stored at Dblk) -> L (lower triangle written back in place)
// and M = L^-1 (full 16 x 16 with explicit zeros above the diagonal, written to Mblk).  Lane y < 16 holds column y of D, lane
// 16 + y column y of the identity; the row operations row_x -= (D_xj / D_jj) row_j, x > j, leave U = diag(p) Lt^T in the D
// part and Lt^-1 (unit lower) in the identity part; L = Lt diag(sqrt p), M = diag(1 / sqrt p) Lt^-1.
// Per pivot the dependent chain is: shuffle (pivot) -> MUFU seed -> 3 DFMA (third-order Newton step, folded into the scaled
// pivot-row entry t) -> DFMA (update) -> next shuffle; the column entries arrive by shuffles that do not depend on the pivot.
// Returns 0 or 1 + index of the first non-positive pivot.
__device__ __forceinline__ int leaf_diag_factor(double *Dblk, double *Mblk, int lane) {
  const int y = lane & 15;
  const bool is_d = lane < 16;
  double v[LB], p[LB];
#pragma unroll
  for (int x = 0; x < LB; ++x) {
    const double dv = (x >= y) ? Dblk[x * LSD + y] : Dblk[y * LSD + x];
    v[x] = is_d ? dv : (x == y ? 1.0 : 0.0);
  }
  int fail = 0;
#pragma unroll
  for (int j = 0; j < LB; ++j) {
    const double pj = shfl_d(v[j], j);
    p[j] = pj;
    if (!(pj > 0.0) && fail == 0) fail = j + 1;
    // t = -v[j] / pj:  r0 ~ 1/pj (20 bits), e = 1 - pj r0, 1/pj = r0 (1 + e + e^2 + O(e^3))
    double r0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(pj));
    const double t0 = -r0 * v[j];
    const double e = fma(-pj, r0, 1.0);
    const double s2 = fma(e, e, e);
    const double t = fma(t0, s2, t0);
#pragma unroll
    for (int x = j + 1; x < LB; ++x) v[x] = fma(shfl_d(v[x], j), t, v[x]);
  }
  double py = p[0];
#pragma unroll
  for (int x = 1; x < LB; ++x) py = (x == y) ? p[x] : py;
  const double rs_own = fast_rsqrt(py);
  // branch-free write-out: lanes < 16 store row y of L (x <= y), lanes >= 16 column y of M
  double *dst = is_d ? Dblk + y * LSD : Mblk + y;
  const int step = is_d ? 1 : LMD;
  const int xmax = is_d ? y : LB;
#pragma unroll
  for (int x = 0; x < LB; ++x) {
    const double val = v[x] * shfl_d(rs_own, x);   // L[y][x] = U[x][y] / sqrt(p_x);  M[x][y] = Lt^-1[x][y] / sqrt(p_x)
    if (x <= xmax) dst[x * step] = val;
  }
  return fail;
}

// Diagonal sub-block k of both results, shared -> global (one warp; row segments of 128 bytes; zeros above the diagonal of L)
__device__ __forceinline__ void leaf_diag_writeout(const double *Dblk, const double *Mblk, double *__restrict__ gL, int lda,
                                                   double *__restrict__ gM, int ldm, int lane) {
  const int y = lane & 15;
#pragma unroll
  for (int it = 0; it < LB / 2; ++it) {
    const int r = 2 * it + (lane >> 4);
    gL[(size_t)r * lda + y] = (y <= r) ? Dblk[r * LSD + y] : 0.0;
    gM[(size_t)r * ldm + y] = Mblk[r * LMD + y];
  }
}

// c (accumulator fragments) -> global 16 x 16 sub-block, 16 bytes per lane
__device__ __forceinline__ void blk_store_global(const double (&c)[2][2][2], double *__restrict__ G, int ld, int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j)
      *reinterpret_cast<double2 *>(G + (size_t)(8 * i + g) * ld + 8 * j + 2 * t) = make_double2(c[i][j][0], c[i][j][1]);
}

// Block row r of M (off-diagonal part):   M_rj = -M_rr T_rj,  T_rj = sum_{m=j}^{r-1} L_rm M_mj.
// leaf_inv_accum adds the terms m_lo <= m < m_hi to the accumulators; leaf_inv_finish multiplies by -M_rr, keeps the result
// transposed in the upper triangle of S (operand of the rows below) and writes it to Mi.
__device__ __forceinline__ void leaf_inv_accum(double (&c)[2][2][2], const double *S, const double *Md, int r, int j, int m_lo, int m_hi,
                                               int lane) {
  for (int m = m_lo; m < m_hi; ++m) {
    if (m == j)
      blk_mma<false, true>(c, S + (r * LB) * LSD + j * LB, LSD, Md + j * LB * LMD, LMD, 1.0, lane);            // M_jj
    else
      blk_mma<false, false>(c, S + (r * LB) * LSD + m * LB, LSD, S + (j * LB) * LSD + m * LB, LSD, 1.0, lane);  // M_mj stored transposed
  }
}
__device__ __forceinline__ void leaf_inv_finish(double (&c)[2][2][2], double *S, const double *Md, double *T, double *__restrict__ Mi,
                                                int ldm, int r, int j, int lane) {
  __syncwarp();
  blk_store(c, T, LMD, lane);
  __syncwarp();
  blk_zero(c);
  blk_mma<false, true>(c, Md + r * LB * LMD, LMD, T, LMD, -1.0, lane);
  blk_store_t(c, S + (j * LB) * LSD + r * LB, LSD, lane);
  blk_store_global(c, Mi + (size_t)(r * LB) * ldm + j * LB, ldm, lane);
}
// tasks j = w, w + nw, ... < r
__device__ __forceinline__ void leaf_inverse_row(double *S, const double *Md, double *T, double *__restrict__ Mi, int ldm, int r, int w,
                                                 int nw, int lane) {
  for (int j = w; j < r; j += nw) {
    double c[2][2][2];
    blk_zero(c);
    leaf_inv_accum(c, S, Md, r, j, j, r, lane);
    leaf_inv_finish(c, S, Md, T, Mi, ldm, r, j, lane);
  }
}

// A_ij -= L_ik L_jk^T for one sub-block
__device__ __forceinline__ void leaf_trailing_block(double *S, int bi, int bj, int k, int lane) {
  double c[2][2][2];
  double *blk = S + (bi * LB) * LSD + bj * LB;
  blk_load(c, blk, LSD, lane);
  blk_mma<false, false>(c, S + (bi * LB) * LSD + k * LB, LSD, S + (bj * LB) * LSD + k * LB, LSD, -1.0, lane);
  blk_store(c, blk, LSD, lane);
}

__device__ __forceinline__ void leaf_cp_async16(double *smem_dst, const double *gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src));
}

// Schedule (look-ahead): after the panel of step k, warp 0 updates the next diagonal sub-block and factors it right away while
// warps 1-7 finish the trailing update of step k and compute block row k of M; one barrier pair per step.  Results leave for
// global memory from the accumulator fragments as they are produced (no epilogue); the zeros above the diagonal are stored
// first and complete under the computation.
// The factorisation proper: S holds the lower triangle of the block (all threads have synchronised after filling it); on return
// (after a __syncthreads) S holds L in its lower triangle and M^T in its strictly-upper sub-blocks, Md the diagonal sub-blocks
// of M, and both results are on their way to A / Mi.  Returns 0 or 1 + index of the first non-positive pivot (valid in warp 0).
__device__ __forceinline__ int leaf_core(double *S, double *Md, double *Tw, double *__restrict__ A, int lda, double *__restrict__ Mi,
                                         int ldm, int tid, int lane, int warp) {
  constexpr int NW = LEAF_THREADS / 32;
  double *T = Tw + warp * LB * LMD;
  int fail = 0;
  double pre[2][2][2];   // warps 1-7: partial sum of their task of the last block row of M, carried across the last barrier
  blk_zero(pre);
  // step k = -1 only factors the first diagonal sub-block (one code site for the factorisation: it stays in the instruction cache)
#pragma unroll 1
  for (int k = -1; k < LNB; ++k) {
    LEAF_CLK(2 + 4 * (k + 1));
    if (k >= 0) {
      // panel: L_ik = A_ik M_kk^T; warp 0 has sub-block k+1 and goes straight on to the next diagonal sub-block, the others
      // wait (named barrier 1) until the whole panel is in shared memory
      for (int i = k + 1 + warp; i < LNB; i += NW) {
        double c[2][2][2];
        blk_zero(c);
        double *blk = S + (i * LB) * LSD + k * LB;
        blk_mma<false, false>(c, blk, LSD, Md + k * LB * LMD, LMD, 1.0, lane);
        __syncwarp();
        blk_store(c, blk, LSD, lane);
        blk_store_global(c, A + (size_t)(i * LB) * lda + k * LB, lda, lane);
      }
      if (warp == 0) {
        __threadfence_block();
        asm volatile("bar.arrive 1, %0;" ::"n"(LEAF_THREADS) : "memory");
      } else {
        asm volatile("bar.sync 1, %0;" ::"n"(LEAF_THREADS) : "memory");
      }
    }
    LEAF_CLK(3 + 4 * (k + 1));
    if (warp == 0) {
      if (k + 1 < LNB) {
        if (k >= 0) {
          __syncwarp();
          leaf_trailing_block(S, k + 1, k + 1, k, lane);
          __syncwarp();
        }
        const int o = (k + 1) * LB;
        const int f = leaf_diag_factor(S + o * LSD + o, Md + (k + 1) * LB * LMD, lane);
        if (f != 0 && fail == 0) fail = o + f;
      }
      LEAF_CLK(4 + 4 * (k + 1));
    } else if (k < 0) {
      // meanwhile: explicit zeros above the diagonal of both outputs, outside the diagonal sub-blocks (leaf_diag_writeout has
      // those); plain stores that complete under the computation
      for (int e = tid - 32; e < TILE * (TILE / 2); e += LEAF_THREADS - 32) {
        const int r = e >> 6, c = (e & 63) * 2;
        if (c > r && (c >> 4) != (r >> 4)) {
          *reinterpret_cast<double2 *>(A + (size_t)r * lda + c) = make_double2(0.0, 0.0);
          *reinterpret_cast<double2 *>(Mi + (size_t)r * ldm + c) = make_double2(0.0, 0.0);
        }
      }
    } else {
      if (warp == 1 + (k % (NW - 1))) {
        const int o = k * LB;
        leaf_diag_writeout(S + o * LSD + o, Md + k * LB * LMD, A + (size_t)o * lda + o, lda, Mi + (size_t)o * ldm + o, ldm, lane);
      }
      // trailing update A_ij -= L_ik L_jk^T, i >= j > k, without the next diagonal sub-block (warp 0 has it)
      const int nt = (LNB - 1 - k) * (LNB - k) / 2;
      for (int t = warp; t < nt; t += NW - 1) {   // t = 0 is sub-block (k+1, k+1)
        int i = 0, rem = t;
        while (rem > i) {
          rem -= i + 1;
          ++i;
        }
        leaf_trailing_block(S, k + 1 + i, k + 1 + rem, k, lane);
      }
      // The last block row of M is the only one without a diagonal factorisation to hide under: warp w owns its task
      // j = LNB - 1 - w and sums the terms m < LNB - 2 one step early (they only need finished rows of M); the pairing evens out
      // the work with this step's row (task w - 1), so that only two products per warp remain after the last factorisation.
      const int jl = LNB - 1 - warp;
      if (k + 1 < LNB) {
        leaf_inverse_row(S, Md, T, Mi, ldm, k, warp - 1, NW - 1, lane);
        if (k + 2 == LNB) {
          blk_zero(pre);
          leaf_inv_accum(pre, S, Md, LNB - 1, jl, jl, LNB - 2, lane);
        }
      } else {
        leaf_inv_accum(pre, S, Md, LNB - 1, jl, jl > LNB - 2 ? jl : LNB - 2, LNB - 1, lane);
        leaf_inv_finish(pre, S, Md, T, Mi, ldm, LNB - 1, jl, lane);
      }
    }
    __syncthreads();
    LEAF_CLK(5 + 4 * (k + 1));
  }
  return fail;
}

__global__ void __launch_bounds__(LEAF_THREADS, 1)
leaf_blocked_kernel(double *__restrict__ A, int lda, double *__restrict__ Mi, int ldm, int index_base, int *info) {
  extern __shared__ double sm[];
  double *S = sm;                          // 128 x 132
  double *Md = S + TILE * LSD;             // 8 x (16 x 20): diagonal sub-blocks of M
  double *Tw = Md + LNB * LB * LMD;        // 8 x (16 x 20): per-warp scratch
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // provably warp-uniform: the role branches below stay convergent
  pdl_trigger();
  pdl_wait();
  LEAF_CLK(0);
  // lower triangle of the block -> shared memory (asynchronous copies, all in flight together)
#pragma unroll
  for (int it = 0; it < TILE * (TILE / 2) / LEAF_THREADS; ++it) {
    const int e = tid + it * LEAF_THREADS, r = e >> 6, c = (e & 63) * 2;
    if (c <= r) leaf_cp_async16(S + r * LSD + c, A + (size_t)r * lda + c);
  }
  asm volatile("cp.async.commit_group;\n" ::);
  asm volatile("cp.async.wait_group 0;\n" ::);
  __syncthreads();
  LEAF_CLK(1);
  const int fail = leaf_core(S, Md, Tw, A, lda, Mi, ldm, tid, lane, warp);
  if (warp == 0 && lane == 0 && fail != 0) atomicCAS(info, 0, index_base + fail);
  LEAF_CLK(40);
  LEAF_CLK(41);
}

// ---------------------------------------------------------------------------------------------------------------------
// N <= 128: the whole NLL + gradient evaluation in ONE kernel (one CTA).  At the sizes of an ordinary BO run (the reference's
// own examples, BASELINE config 1) an evaluation through the general path is eleven launches of a few microseconds each;
// here the covariance block is built straight into the shared array of the blocked leaf, factorised and inverted in place,
// and alpha = M^T (M y), log det, alpha.y, Ky^-1 = M^T M and the D + 2 gradient sums follow without leaving the SM.
// Same formulas and outputs as kmat_kernel (mode 3) -> leaf -> trmv -> logdet / dot -> potri -> kgrad_kernel<.., FUSED>:
//   scal[0] = log det, scal[1] = alpha.y, scal[2] = sum K.G, scal[3] = tr G, scal[4 + q] = sum (k'/r) G ds_q^2,  G = (alpha alpha^T - W) / 2
// ---------------------------------------------------------------------------------------------------------------------
constexpr int TINY_DMAX = 32;

struct TinyLs {
  double ls[TINY_DMAX];   // per-dimension lengthscales, by value: no upload, no separate scaling kernel
};

template <int KIND>
__global__ void __launch_bounds__(LEAF_THREADS, 1)
tiny_fit_kernel(const double *__restrict__ X, TinyLs lsv, double *__restrict__ XsT, double *__restrict__ ls_dev, double *__restrict__ inv_ls_dev,
                int n, int d, double variance, double diag_add, const double *__restrict__ y, double *__restrict__ A,
                double *__restrict__ Mi, double *__restrict__ W, int ld, double *__restrict__ z_out, double *__restrict__ alpha_out,
                double *__restrict__ scal, int want_grad, int *info) {
  extern __shared__ double sm[];
  double *S = sm;
  double *Md = S + TILE * LSD;
  double *Tw = Md + LNB * LB * LMD;
  double *xs = Tw + LNB * LB * LMD;        // d x 128 scaled inputs, dimension-major
  double *yv = xs + TINY_DMAX * TILE;      // 128
  double *zv = yv + TILE;                  // 128
  double *av = zv + TILE;                  // 128
  double *red = av + TILE;                 // 8 x (TINY_DMAX + 2)
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  constexpr int NW = LEAF_THREADS / 32;
  pdl_trigger();
  pdl_wait();
  // scaled inputs, dimension-major (scale_transpose_kernel): kept for the predictive calls that follow
  for (int e = tid; e < d * TILE; e += LEAF_THREADS) {
    const int q = e >> 7, i = e & 127;
    double v = (i < n) ? X[(size_t)i * d + q] / lsv.ls[q] : 0.0;
    v = v > 1e150 ? 1e150 : (v < -1e150 ? -1e150 : v);
    xs[e] = v;
    XsT[e] = v;
  }
  if (tid < d) {
    ls_dev[tid] = lsv.ls[tid];
    inv_ls_dev[tid] = 1.0 / lsv.ls[tid];
  }
  if (tid < TILE) yv[tid] = (tid < n) ? y[tid] : 0.0;
  __syncthreads();
  // ---- Ky = K + diag_add I, identity padding (exact_gaussian_inference.py:55-56), lower triangle ----
  // 16 x 16 threads, 8 x 8 register tile: thread (ty, tx) owns the pairs (ty + 16 a, tx + 16 b), b <= a (like kmat_kernel)
  const int tx = tid & 15, ty = tid >> 4;
  {
    double r2[LNB][LNB];
#pragma unroll
    for (int a2 = 0; a2 < LNB; ++a2)
#pragma unroll
      for (int b2 = 0; b2 < LNB; ++b2) r2[a2][b2] = 0.0;
    for (int q = 0; q < d; ++q) {
      double va[LNB], vb[LNB];
#pragma unroll
      for (int a2 = 0; a2 < LNB; ++a2) {
        va[a2] = xs[q * TILE + ty + LB * a2];
        vb[a2] = xs[q * TILE + tx + LB * a2];
      }
#pragma unroll
      for (int a2 = 0; a2 < LNB; ++a2)
#pragma unroll
        for (int b2 = 0; b2 <= a2; ++b2) {
          const double df = va[a2] - vb[b2];
          r2[a2][b2] = fma(df, df, r2[a2][b2]);
        }
    }
#pragma unroll
    for (int a2 = 0; a2 < LNB; ++a2)
#pragma unroll
      for (int b2 = 0; b2 <= a2; ++b2) {
        const int i = ty + LB * a2, j = tx + LB * b2;
        if (j > i) continue;
        double v;
        if (i < n) {
          v = cov_k<KIND>(r2[a2][b2], variance);
          if (i == j) v += diag_add;
        } else {
          v = (i == j) ? 1.0 : 0.0;
        }
        S[i * LSD + j] = v;
      }
  }
  __syncthreads();
  const int fail = leaf_core(S, Md, Tw, A, ld, Mi, ld, tid, lane, warp);
  if (warp == 0 && lane == 0) {
    *info = fail;
    scal[d + 4] = (double)fail;   // travels with the results: one device-to-host copy per evaluation
  }
  // ---- z = M y, alpha = M^T z (dpotrs, linalg.py:116-125); M[r][c] = S[c][r] outside the diagonal sub-blocks, Md inside ----
  if (tid < TILE) {
    const int r = tid, rb = r >> 4;
    double acc = 0.0;
    for (int c = 0; c < rb * LB; ++c) acc = fma(S[c * LSD + r], yv[c], acc);
    for (int c = rb * LB; c <= r; ++c) acc = fma(Md[rb * LB * LMD + (r & 15) * LMD + (c & 15)], yv[c], acc);
    zv[r] = acc;
  }
  __syncthreads();
  if (tid < TILE) {
    const int c = tid, cb = c >> 4;
    double acc = 0.0;
    for (int r = c; r < (cb + 1) * LB; ++r) acc = fma(Md[cb * LB * LMD + (r & 15) * LMD + (c & 15)], zv[r], acc);
    for (int r = (cb + 1) * LB; r < TILE; ++r) acc = fma(S[c * LSD + r], zv[r], acc);
    av[c] = acc;
    alpha_out[c] = acc;
    z_out[c] = zv[c];
  } else {
    // log det = 2 sum log L_ii and alpha.y come from the second half of the block
    const int t = tid - TILE;
    double ld_acc = (t < n) ? log(S[t * LSD + t]) : 0.0;
    ld_acc = warp_sum(ld_acc);
    if (lane == 0) red[warp] = ld_acc;
  }
  __syncthreads();
  if (tid == 0) scal[0] = 2.0 * ((red[4] + red[5]) + (red[6] + red[7]));
  if (warp == 1) {
    double acc = 0.0;
    for (int i = lane; i < TILE; i += 32) acc = fma(av[i], yv[i], acc);
    acc = warp_sum(acc);
    if (lane == 0) scal[1] = acc;
  }
  if (!want_grad) return;
  __syncthreads();   // everyone is done with L (log det) before W overwrites the lower triangle
  // ---- W = Ky^-1 = M^T M (dpotri): sub-block (i, j), i >= j:  sum_{k >= i} M_ki^T M_kj; kept in S (lower) and written to W in full ----
  for (int t = warp; t < LNB * (LNB + 1) / 2; t += NW) {
    int i = 0, rem = t;
    while (rem > i) {
      rem -= i + 1;
      ++i;
    }
    const int j = rem;
    double c[2][2][2];
    blk_zero(c);
    // k = i: M_ii^T from Md; M_ij from the upper triangle (or Md when j == i)
    if (j == i)
      blk_mma<true, true>(c, Md + i * LB * LMD, LMD, Md + i * LB * LMD, LMD, 1.0, lane);
    else
      blk_mma<true, false>(c, Md + i * LB * LMD, LMD, S + (j * LB) * LSD + i * LB, LSD, 1.0, lane);
    for (int k = i + 1; k < LNB; ++k)
      blk_mma<false, false>(c, S + (i * LB) * LSD + k * LB, LSD, S + (j * LB) * LSD + k * LB, LSD, 1.0, lane);
    // the products only read the upper triangle and Md, so the lower triangle (L, already in A) can take W
    blk_store(c, S + (i * LB) * LSD + j * LB, LSD, lane);
    blk_store_global(c, W + (size_t)(i * LB) * ld + j * LB, ld, lane);
    if (j != i) {
      const int g = lane >> 2, tq = lane & 3;
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          W[(size_t)(j * LB + 8 * b + 2 * tq) * ld + i * LB + 8 * a + g] = c[a][b][0];
          W[(size_t)(j * LB + 8 * b + 2 * tq + 1) * ld + i * LB + 8 * a + g] = c[a][b][1];
        }
    }
  }
  __syncthreads();
  // ---- gradient sums over all pairs (update_gradients_full, stationary.py:218-238, with dL_dK = (alpha alpha^T - W) / 2) ----
  // same register tiling as the build; the pair weights replace r2, then one pass per dimension (like kgrad_kernel)
  double r2[LNB][LNB];
#pragma unroll
  for (int a2 = 0; a2 < LNB; ++a2)
#pragma unroll
    for (int b2 = 0; b2 < LNB; ++b2) r2[a2][b2] = 0.0;
  for (int q = 0; q < d; ++q) {
    double va[LNB], vb[LNB];
#pragma unroll
    for (int a2 = 0; a2 < LNB; ++a2) {
      va[a2] = xs[q * TILE + ty + LB * a2];
      vb[a2] = xs[q * TILE + tx + LB * a2];
    }
#pragma unroll
    for (int a2 = 0; a2 < LNB; ++a2)
#pragma unroll
      for (int b2 = 0; b2 <= a2; ++b2) {
        const double df = va[a2] - vb[b2];
        r2[a2][b2] = fma(df, df, r2[a2][b2]);
      }
  }
  double acc_var = 0.0, acc_tr = 0.0;
#pragma unroll
  for (int a2 = 0; a2 < LNB; ++a2)
#pragma unroll
    for (int b2 = 0; b2 <= a2; ++b2) {
      const int i = ty + LB * a2, j = tx + LB * b2;
      double wgt = 0.0;
      if (j <= i && i < n) {
        const double g = 0.5 * (av[i] * av[j] - S[i * LSD + j]);
        const double wt = (i == j) ? 1.0 : 2.0;
        if (i == j) acc_tr += g;
        double k, dk;
        cov_k_dk<KIND>(r2[a2][b2], variance, k, dk);
        acc_var = fma(wt * k, g, acc_var);
        wgt = wt * dk * g;
      }
      r2[a2][b2] = wgt;
    }
  acc_var = warp_sum(acc_var);
  acc_tr = warp_sum(acc_tr);
  if (lane == 0) {
    red[warp * (TINY_DMAX + 2) + 0] = acc_var;
    red[warp * (TINY_DMAX + 2) + 1] = acc_tr;
  }
  for (int q = 0; q < d; ++q) {
    double va[LNB], vb[LNB];
#pragma unroll
    for (int a2 = 0; a2 < LNB; ++a2) {
      va[a2] = xs[q * TILE + ty + LB * a2];
      vb[a2] = xs[q * TILE + tx + LB * a2];
    }
    double acc4[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int a2 = 0; a2 < LNB; ++a2)
#pragma unroll
      for (int b2 = 0; b2 <= a2; ++b2) {
        const double df = va[a2] - vb[b2];
        acc4[a2 & 3] = fma(r2[a2][b2], df * df, acc4[a2 & 3]);
      }
    const double v = warp_sum((acc4[0] + acc4[1]) + (acc4[2] + acc4[3]));
    if (lane == 0) red[warp * (TINY_DMAX + 2) + 2 + q] = v;
  }
  __syncthreads();
  if (tid < d + 2) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < NW; ++w) v += red[w * (TINY_DMAX + 2) + tid];
    scal[2 + tid] = v;
  }
}

// want_grad: also W = Ky^-1 and the gradient sums.  X: n x d raw inputs; ls: d lengthscales (host); XsT (d x 128), ls_dev, inv_ls_dev:
// outputs for the predictive calls; y, z, alpha: 128-vectors; scal: d + 5 doubles (the last one = info).
int launch_tiny_fit(int kind, const double *X, const double *ls_host, double *XsT, double *ls_dev, double *inv_ls_dev, int n, int d,
                    double variance, double diag_add, const double *y, Factor &f, double *z, double *alpha, double *scal, int want_grad) {
  GPB_REQUIRE(f.np == TILE && n <= TILE && d <= TINY_DMAX, "tiny fit: needs N <= 128 and D <= %d", TINY_DMAX);
  const size_t smem = (size_t)(TILE * LSD + 2 * LNB * LB * LMD + TINY_DMAX * TILE + 3 * TILE + (LEAF_THREADS / 32) * (TINY_DMAX + 2)) * sizeof(double);
  static FuncConfigMask configured{0};
  FuncConfigOnce once_configured(configured);
  if (once_configured.needed) {
    GPB_CUDA(cudaFuncSetAttribute(tiny_fit_kernel<GPB_KERN_RBF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GPB_CUDA(cudaFuncSetAttribute(tiny_fit_kernel<GPB_KERN_MATERN52>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  TinyLs lsv;
  for (int q = 0; q < TINY_DMAX; ++q) lsv.ls[q] = q < d ? ls_host[q] : 1.0;
  if (kind == GPB_KERN_RBF)
    GPB_CUDA(launch_pdl(tiny_fit_kernel<GPB_KERN_RBF>, dim3(1), dim3(LEAF_THREADS), smem, f.stream, X, lsv, XsT, ls_dev, inv_ls_dev, n, d,
                        variance, diag_add, y, f.A, f.Mi, f.W, TILE, z, alpha, scal, want_grad, f.info));
  else
    GPB_CUDA(launch_pdl(tiny_fit_kernel<GPB_KERN_MATERN52>, dim3(1), dim3(LEAF_THREADS), smem, f.stream, X, lsv, XsT, ls_dev, inv_ls_dev, n,
                        d, variance, diag_add, y, f.A, f.Mi, f.W, TILE, z, alpha, scal, want_grad, f.info));
  count_launch();
  f.l_pending = false;
  f.l_from = 0;
  return 0;
}

// =====================================================================================================================
// Cooperative diagonal-block solver: cholinv of a block of up to COOP_MAX rows as ONE persistent kernel
// =====================================================================================================================
// MEASURED AND SWITCHED OFF (GPB_COOP_N=1024 selects it; default 0).  The idea (round-1 review, item 5): below ~1024 rows the
// recursion is a chain of dependent launches that each last 6 - 15 us (84 of them on the critical path of an N = 4096 evaluation,
// profiles/r2i_launches_nll4096.md); one grid of co-resident CTAs could walk the same recursion itself -- explicit stack, the same
// order of operations -- and separate dependent steps by grid-wide barriers instead of kernel boundaries.  Built, bit-identical to
// the launch chain (tests/test_gpu_native.py), and SLOWER: N = 4096 4.32 ms against 3.70, N = 1024 0.535 against 0.437, N = 256
// 0.189 against 0.158 (profiles/r2q_coop_solver.json).  A 32 x 32 tile's k-loop is a latency chain (16 dependent DMMA issue slots
// per 16 k); the stand-alone launches hide it with 16 warps per SM and programmatic dependent launch, a persistent grid that must
// also hold the leaf's 176 KB of shared memory runs 8 warps per SM, two tiles at a time -- its product steps take about twice as
// long as the launches they replace, which the cheaper barriers do not buy back.  A leaf is executed by CTA 0 with the code of leaf_blocked_kernel; a product step
// is cut into the 32 x 32 tiles of the stand-alone engine's small configuration (BK 16, 3 stages, the same fragment order, so every
// output element is the bit pattern gemm_dmma_kernel<.., 32, 32, ..> produces), dealt to the two 4-warp halves of every CTA;
// T12 = M11^T L21^T and A22 -= L21 L21^T depend on L21 only and share a step.
namespace cgc = cooperative_groups;

struct CoopArgs {
  double *A, *Mi, *W;
  int ld, off, n;
  int *info;
};

__device__ __forceinline__ void half_bar(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }

// one 32 x 32 tile of C = alpha op(A) op(B)^T + beta C, executed by 128 threads (tid in [0, 128)) with their own staging area
template <int LA, int LB>
__device__ __forceinline__ void coop_gemm_tile(const GemmArgs &p, int t, double *smem, int tid, int bar_id) {
  constexpr int BM = 32, BN = 32, BK = 16, STAGES = 3, THREADS = 128, WARPS_N = 2;
  constexpr int WTM = 16, WTN = 16, MI = 2, NI = 2;
  constexpr int A_TILE = tile_doubles<LA, BM, BK>();
  constexpr int B_TILE = tile_doubles<LB, BN, BK>();
  double *sA = smem, *sB = smem + STAGES * A_TILE;
  int tm, tn;
  if (p.tri_out) {
    const int q = t >> 4, w_in = t & 15;                           // 16 tiles per 128-block; lower blocks in row-major order
    int i = (int)((sqrt(8.0 * (double)q + 1.0) - 1.0) * 0.5);
    while ((i + 1) * (i + 2) / 2 <= q) ++i;
    while (i * (i + 1) / 2 > q) --i;
    const int j = q - i * (i + 1) / 2;
    tm = i * 4 + (w_in >> 2);
    tn = j * 4 + (w_in & 3);
  } else {
    const int tiles_n = p.N / BN;
    tm = t / tiles_n;
    tn = t - tm * tiles_n;
  }
  const int row0 = tm * BM, col0 = tn * BN;
  const int rblk = row0 & ~127, cblk = col0 & ~127;
  const int klo = (p.klo_mode == 1) ? rblk : (p.klo_mode == 2) ? cblk : 0;
  int khi = (p.khi_mode == 1) ? rblk + 128 : (p.khi_mode == 2) ? cblk + 128 : p.K;
  if (khi > p.K) khi = p.K;
  const int ktiles = (khi - klo) / BK;
  const int warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, tq = lane & 3;
  const int wm = (warp / WARPS_N) * WTM, wn = (warp % WARPS_N) * WTN;
  const double *Abase = (LA == LAYOUT_ROWK) ? p.A + (size_t)row0 * p.lda : p.A + row0;
  const double *Bbase = (LB == LAYOUT_ROWK) ? p.B + (size_t)col0 * p.ldb : p.B + col0;
  double acc[MI][NI][2];
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < ktiles) {
      load_tile<LA, BM, THREADS, BK>(sA + s * A_TILE, Abase, p.lda, klo + s * BK, tid);
      load_tile<LB, BN, THREADS, BK>(sB + s * B_TILE, Bbase, p.ldb, klo + s * BK, tid);
    }
    cp_async_commit();
  }
  for (int kt = 0; kt < ktiles; ++kt) {
    cp_async_wait<STAGES - 2>();
    half_bar(bar_id);
    {
      const int nt = kt + STAGES - 1;
      if (nt < ktiles) {
        const int s = nt % STAGES;
        load_tile<LA, BM, THREADS, BK>(sA + s * A_TILE, Abase, p.lda, klo + nt * BK, tid);
        load_tile<LB, BN, THREADS, BK>(sB + s * B_TILE, Bbase, p.ldb, klo + nt * BK, tid);
      }
      cp_async_commit();
    }
    const double *a_s = sA + (kt % STAGES) * A_TILE;
    const double *b_s = sB + (kt % STAGES) * B_TILE;
#pragma unroll
    for (int ks = 0; ks < BK / 4; ++ks) {
      const int k0 = ks * 4 + tq;
      double a[MI], b[NI];
#pragma unroll
      for (int i = 0; i < MI; ++i) a[i] = frag<LA, BM, BK>(a_s, wm + i * 8 + g, k0);
#pragma unroll
      for (int j = 0; j < NI; ++j) b[j] = frag<LB, BN, BK>(b_s, wn + j * 8 + g, k0);
#pragma unroll
      for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
  }
  cp_async_wait<0>();
  const double alpha = p.alpha, beta = p.beta;
#pragma unroll
  for (int i = 0; i < MI; ++i) {
    const int row = row0 + wm + i * 8 + g;
#pragma unroll
    for (int j = 0; j < NI; ++j) {
      const int col = col0 + wn + j * 8 + 2 * tq;
      double2 *ptr = reinterpret_cast<double2 *>(p.C + (size_t)row * p.ldc + col);
      double2 o;
      o.x = alpha * acc[i][j][0];
      o.y = alpha * acc[i][j][1];
      if (beta != 0.0) {
        const double2 c = *ptr;
        o.x += beta * c.x;
        o.y += beta * c.y;
      }
      *ptr = o;
    }
  }
  half_bar(bar_id);            // the staging area is free for the half's next tile
}

__device__ __forceinline__ int coop_tiles(const GemmArgs &p) {
  const int bm = p.M / 128, bn = p.N / 128;
  return (p.tri_out ? bm * (bm + 1) / 2 : bm * bn) * 16;
}

// one step: up to two independent products (la / lb: storage orders of the first one; the second is always ROWK x ROWK)
__device__ __forceinline__ void coop_step(const GemmArgs &g0, int la0, const GemmArgs *g1, double *smem_half, int tid128, int half,
                                          int G) {
  const int t0 = coop_tiles(g0), t1 = g1 ? coop_tiles(*g1) : 0;
  for (int w = blockIdx.x * 2 + half; w < t0 + t1; w += 2 * G) {
    if (w < t0) {
      if (la0 == LAYOUT_COLK) coop_gemm_tile<LAYOUT_COLK, LAYOUT_ROWK>(g0, w, smem_half, tid128, 1 + half);
      else coop_gemm_tile<LAYOUT_ROWK, LAYOUT_ROWK>(g0, w, smem_half, tid128, 1 + half);
    } else {
      coop_gemm_tile<LAYOUT_ROWK, LAYOUT_ROWK>(*g1, w - t0, smem_half, tid128, 1 + half);
    }
  }
}

constexpr int COOP_GEMM_SMEM = 3 * (32 * 20 + 32 * 20);   // doubles per half (the larger of the two layouts)

__global__ void __launch_bounds__(LEAF_THREADS, 1) cholinv_coop_kernel(const CoopArgs a) {
  cgc::grid_group grid = cgc::this_grid();
  extern __shared__ double sm[];
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int half = warp >> 2, tid128 = tid & 127;
  const int G = gridDim.x, ld = a.ld;
  double *smem_half = sm + half * COOP_GEMM_SMEM;
  int s_off[8], s_n[8], s_state[8];
  int sp = 0;
  s_off[0] = a.off;
  s_n[0] = a.n;
  s_state[0] = 0;
  while (sp >= 0) {
    const int off = s_off[sp], n = s_n[sp];
    if (n == TILE) {
      if (blockIdx.x == 0) {                       // the leaf: exactly the body of leaf_blocked_kernel
        double *S = sm, *Md = S + TILE * LSD, *Tw = Md + LNB * LB * LMD;
        double *Ab = a.A + (size_t)off * ld + off, *Mb = a.Mi + (size_t)off * ld + off;
#pragma unroll
        for (int it = 0; it < TILE * (TILE / 2) / LEAF_THREADS; ++it) {
          const int e = tid + it * LEAF_THREADS, r = e >> 6, c = (e & 63) * 2;
          if (c <= r) leaf_cp_async16(S + r * LSD + c, Ab + (size_t)r * ld + c);
        }
        asm volatile("cp.async.commit_group;\n" ::);
        asm volatile("cp.async.wait_group 0;\n" ::);
        __syncthreads();
        const int fail = leaf_core(S, Md, Tw, Ab, ld, Mb, ld, tid, lane, warp);
        if (warp == 0 && lane == 0 && fail != 0) atomicCAS(a.info, 0, off + fail);
        __syncthreads();
      }
      grid.sync();
      --sp;
      continue;
    }
    const int h = ((n / TILE) / 2) * TILE, r = n - h;
    double *A21 = a.A + (size_t)(off + h) * ld + off;
    double *A22 = a.A + (size_t)(off + h) * ld + off + h;
    double *M11 = a.Mi + (size_t)off * ld + off;
    double *M21 = a.Mi + (size_t)(off + h) * ld + off;
    double *M22 = a.Mi + (size_t)(off + h) * ld + off + h;
    double *S21 = a.W + (size_t)(off + h) * ld + off;
    double *T12 = a.W + (size_t)off * ld + off + h;
    if (s_state[sp] == 0) {                        // cholinv(A11)
      s_state[sp] = 1;
      ++sp;
      s_off[sp] = off;
      s_n[sp] = h;
      s_state[sp] = 0;
    } else if (s_state[sp] == 1) {
      // L21 = A21 M11^T
      const GemmArgs g1{A21, ld, M11, ld, S21, ld, r, h, h, 1.0, 0.0, 0, 0, 2};
      coop_step(g1, LAYOUT_ROWK, nullptr, smem_half, tid128, half, G);
      grid.sync();
      // T12 = M11^T L21^T   and   A22 -= L21 L21^T
      const GemmArgs g2{M11, ld, S21, ld, T12, ld, h, r, h, 1.0, 0.0, 0, 1, 0};
      const GemmArgs g3{S21, ld, S21, ld, A22, ld, r, r, h, -1.0, 1.0, 1, 0, 0};
      coop_step(g2, LAYOUT_COLK, &g3, smem_half, tid128, half, G);
      grid.sync();
      s_state[sp] = 2;                             // cholinv(A22)
      ++sp;
      s_off[sp] = off + h;
      s_n[sp] = r;
      s_state[sp] = 0;
    } else {
      // M21 = -M22 T21
      const GemmArgs g4{M22, ld, T12, ld, M21, ld, r, h, r, -1.0, 0.0, 0, 0, 1};
      coop_step(g4, LAYOUT_ROWK, nullptr, smem_half, tid128, half, G);
      grid.sync();
      --sp;
    }
  }
}

static int coop_max_n() {
  static const int v = [] {
    const char *e = getenv("GPB_COOP_N");
    return e ? atoi(e) : 0;
  }();
  return v;
}

// cholinv of the diagonal block [off, off + n) in one cooperative launch (n a multiple of 128, 256 <= n <= coop_max_n())
static int launch_cholinv_coop(Factor &f, int off, int n) {
  const size_t smem_b = (size_t)(TILE * LSD + 2 * LNB * LB * LMD) * sizeof(double);
  static std::atomic<int> sms[64];
  static FuncConfigMask configured{0};
  int dev = 0;
  GPB_CUDA(cudaGetDevice(&dev));
  GPB_REQUIRE(dev >= 0 && dev < 64, "cholinv_coop: device ordinal %d out of range", dev);
  {
    FuncConfigOnce once(configured);
    if (once.needed) {
      int nsm = 0, occ = 0;
      GPB_CUDA(cudaFuncSetAttribute(cholinv_coop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b));
      GPB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, cholinv_coop_kernel, LEAF_THREADS, smem_b));
      GPB_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
      GPB_REQUIRE(occ >= 1 && nsm >= 1, "cholinv_coop: kernel does not fit an SM");
      sms[dev].store(nsm, std::memory_order_release);
    }
  }
  // never more CTAs than the largest step has half-tiles for (a step of an n-row block has at most 2 (n / 64)^2 / 4 ... tiles)
  const int h = ((n / TILE) / 2) * TILE, r = n - h;
  const int max_tiles = ((h / 128) * (r / 128) + (r / 128) * (r / 128 + 1) / 2) * 16;
  const int G = std::max(1, std::min(sms[dev].load(), (max_tiles + 1) / 2));
  CoopArgs a{f.A, f.Mi, f.W, f.np, off, n, f.info};
  void *args[] = {&a};
  GPB_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void *>(cholinv_coop_kernel), dim3(G), dim3(LEAF_THREADS), args, smem_b, f.stream));
  count_launch();
  return 0;
}

static int g_leaf_variant = -1;  // 1: blocked DMMA leaf (default), 0: register rank-1 sweep (GPB_LEAF=0)

static int launch_leaf(Factor &f, int off, int mode) {
  if (g_leaf_variant < 0) {
    const char *e = getenv("GPB_LEAF");
    g_leaf_variant = (e && e[0] == '0') ? 0 : 1;
  }
  if (mode == 0 && g_leaf_variant == 1) {
    const size_t smem_b = (size_t)(TILE * LSD + 2 * LNB * LB * LMD) * sizeof(double);
    static FuncConfigMask configured_b{0};
    FuncConfigOnce once_configured_b(configured_b);
    if (once_configured_b.needed)
      GPB_CUDA(cudaFuncSetAttribute(leaf_blocked_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b));
    GPB_CUDA(launch_pdl(leaf_blocked_kernel, dim3(1), dim3(LEAF_THREADS), smem_b, f.stream, f.A + (size_t)off * f.np + off, f.np,
                        f.Mi + (size_t)off * f.np + off, f.np, off, f.info));
    count_launch();
    return 0;
  }
  const size_t smem = (size_t)(2 + 3 * TILE + TILE * LEAF_LD) * sizeof(double);
  static FuncConfigMask configured{0};
  FuncConfigOnce once_configured(configured);
  if (once_configured.needed) {
    GPB_CUDA(cudaFuncSetAttribute(leaf_potrf_inv_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GPB_CUDA(cudaFuncSetAttribute(leaf_potrf_inv_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
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

// cooperative solver usable for this call?  (blocked leaf variant; the stream is not being captured into a CUDA graph)
static bool coop_usable(const Factor &f) {
  if (coop_max_n() < 2 * TILE) return false;
  if (g_leaf_variant < 0) {
    const char *e = getenv("GPB_LEAF");
    g_leaf_variant = (e && e[0] == '0') ? 0 : 1;
  }
  if (g_leaf_variant != 1) return false;
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(f.stream, &st) != cudaSuccess) {
    (void)cudaGetLastError();
    return false;
  }
  return st == cudaStreamCaptureStatusNone;
}

// ---------------------------------------------------------------------------------------------------------------------
// potrf + triangular inverse (recursive)
// ---------------------------------------------------------------------------------------------------------------------
// Scratch layout inside W during the factorisation (every 128-block of W is used by exactly one recursion level):
//   strictly-lower block (i, j), i > j : L21 of the level whose (2,1) quadrant contains it  (S21 = A21 M11^T)
//   strictly-upper block (i, j), i < j : T21^T of that level                               (T12 = M11^T L21^T)
// Nothing on the critical path reads L back from A (a parent level only needs M11 and the untouched rows of Ky below), so
// L21 is NOT copied into A level by level: factor_finalize_L moves all strictly-lower blocks in one pass before W is reused
// (Ky^-1 = M^T M) or when the caller asks for L.
//
// Two-stream schedule (f.ov != nullptr): T12 only needs L21 and M11, so it is issued on a low-priority side stream (one per
// recursion depth) right after L21 exists and runs underneath the critical path
//     A22 -= L21 L21^T  ->  cholinv(A22)
// whose lower levels (leaves, 32x32-tile GEMMs) leave most of the 148 SMs idle; M21 = -M22 T21 joins the two streams.
// f.stream must be a high-priority stream so that the small critical-path kernels get the next free SM slots while a long
// side GEMM is resident.
// split > 0 (top level of an append): the leading split x split block is already factorised and inverted -- only the
// trailing block rows are (re)computed: O(N^2 r) instead of O(N^3).
static int cholinv(Factor &f, int off, int n, int depth, int split = 0) {
  if (n == TILE) return launch_leaf(f, off, 0);
  // small diagonal blocks: the whole sub-recursion in one persistent cooperative kernel (not under stream capture: cooperative
  // launches cannot be captured; not with the register leaf variant or the int8 engine on these rows)
  if (split == 0 && n <= coop_max_n() && n >= 2 * TILE && coop_usable(f) && !(ozaki_min_n() > 0 && n >= ozaki_min_n()))
    return launch_cholinv_coop(f, off, n);
  const int ld = f.np;
  const int h = split > 0 ? split : ((n / TILE) / 2) * TILE;  // first part (multiple of 128); default: half, n - h >= h
  const int r = n - h;
  if (split == 0) GPB_TRY(cholinv(f, off, h, depth + 1));
  double *A21 = f.A + (size_t)(off + h) * ld + off;
  double *A22 = f.A + (size_t)(off + h) * ld + off + h;
  double *M11 = f.Mi + (size_t)off * ld + off;
  double *M21 = f.Mi + (size_t)(off + h) * ld + off;
  double *M22 = f.Mi + (size_t)(off + h) * ld + off + h;
  double *S21 = f.W + (size_t)(off + h) * ld + off;   // r x h: L21
  double *T12 = f.W + (size_t)off * ld + off + h;     // h x r: T21^T
  // n >= ozaki_min_n(): the four products of this level run on the int8 tensor cores (gpb_ozaki.cu; experimental, off by default).
  // The engine keeps one plane workspace per (device, stream), so the T12 product of such a level can fork to the side stream like
  // the DMMA one: its residue extraction, MMAs and reconstruction then run underneath A22 -= L21 L21^T and the recursion into A22
  // (GPB_OZAKI_FORK=0 restores the serial order of round 1).
  const bool oz = split == 0 && ozaki_min_n() > 0 && n >= ozaki_min_n();   // (an append keeps the DMMA engine: r is small)
  static const bool oz_fork = [] { const char *e = getenv("GPB_OZAKI_FORK"); return !e || atoi(e) != 0; }();
  const bool fork = (!oz || oz_fork) && f.ov != nullptr && depth < FactorOverlap::MAX_DEPTH && n >= f.ov->min_n;
  GemmArgs g;
  // L21 = A21 * M11^T      (M11 lower: k <= column tile)
  g = GemmArgs{A21, ld, M11, ld, S21, ld, r, h, h, 1.0, 0.0, 0, 0, 2};
  if (oz) GPB_TRY(ozaki_gemm_launch(LAYOUT_ROWK, LAYOUT_ROWK, g, 0, 1, 0, f.stream));
  else GPB_TRY(gemm_launch(LAYOUT_ROWK, LAYOUT_ROWK, g, f.stream));
  cudaStream_t s5 = f.stream;
  if (fork) {
    s5 = f.ov->side[depth];
    GPB_CUDA(cudaEventRecord(f.ov->fork[depth], f.stream));
    GPB_CUDA(cudaStreamWaitEvent(s5, f.ov->fork[depth], 0));
  }
  // T12 = M11^T * L21^T    (M11 lower: k >= row tile of T12)
  g = GemmArgs{M11, ld, S21, ld, T12, ld, h, r, h, 1.0, 0.0, 0, 1, 0};
  if (oz) GPB_TRY(ozaki_gemm_launch(LAYOUT_COLK, LAYOUT_ROWK, g, 2, 0, 0, s5));
  else GPB_TRY(gemm_launch(LAYOUT_COLK, LAYOUT_ROWK, g, s5));
  if (fork) GPB_CUDA(cudaEventRecord(f.ov->join[depth], s5));
  // A22 -= L21 * L21^T     (lower tiles)
  g = GemmArgs{S21, ld, S21, ld, A22, ld, r, r, h, -1.0, 1.0, 1, 0, 0};
  if (oz) GPB_TRY(ozaki_gemm_launch(LAYOUT_ROWK, LAYOUT_ROWK, g, 0, 0, 0, f.stream));
  else GPB_TRY(gemm_launch(LAYOUT_ROWK, LAYOUT_ROWK, g, f.stream));
  GPB_TRY(cholinv(f, off + h, r, depth + 1));
  if (fork) GPB_CUDA(cudaStreamWaitEvent(f.stream, f.ov->join[depth], 0));
  // M21 = -M22 * T21       (M22 lower: k <= row tile; T21 read through its transpose T12, k-contiguous)
  g = GemmArgs{M22, ld, T12, ld, M21, ld, r, h, r, -1.0, 0.0, 0, 0, 1};
  if (oz) GPB_TRY(ozaki_gemm_launch(LAYOUT_ROWK, LAYOUT_ROWK, g, 1, 0, 0, f.stream));
  else GPB_TRY(gemm_launch(LAYOUT_ROWK, LAYOUT_ROWK, g, f.stream));
  return 0;
}

// dst block (bi, bj) = src block (bi, bj) for the 128-blocks strictly below the diagonal in block rows bi >= bi0
__global__ void copy_lower_blocks_kernel(double *__restrict__ dst, const double *__restrict__ src, int ld, int nb, int t0) {
  const int t = blockIdx.x + t0;  // strictly-lower block index: t = bi (bi - 1) / 2 + bj, bi >= 1
  int bi = (int)((sqrt(8.0 * (double)t + 1.0) + 1.0) * 0.5);
  while (bi * (bi - 1) / 2 > t) --bi;
  while ((bi + 1) * bi / 2 <= t) ++bi;
  const int bj = t - bi * (bi - 1) / 2;
  if (bi >= nb) return;
  const size_t base = (size_t)(bi * TILE) * ld + bj * TILE;
  for (int e = threadIdx.x; e < TILE * TILE / 2; e += blockDim.x) {
    const int rrow = e >> 6, c2 = e & 63;
    reinterpret_cast<double2 *>(dst + base + (size_t)rrow * ld)[c2] = reinterpret_cast<const double2 *>(src + base + (size_t)rrow * ld)[c2];
  }
}

// Move the off-diagonal blocks of L from their scratch place in W into A (no-op when already done).
int factor_finalize_L(Factor &f) {
  if (!f.l_pending) return 0;
  const int nb = f.np / TILE;
  const int t0 = f.l_from * (f.l_from - 1) / 2;          // blocks of the block rows < l_from are already in place
  const int blocks = nb * (nb - 1) / 2 - (t0 > 0 ? t0 : 0);
  if (blocks > 0) {
    copy_lower_blocks_kernel<<<blocks, 256, 0, f.stream>>>(f.A, f.W, f.np, nb, t0 > 0 ? t0 : 0);
    count_launch();
    GPB_CHECK_LAUNCH();
  }
  f.l_pending = false;
  return 0;
}

int factor_overlap_create(FactorOverlap **out) {
  FactorOverlap *ov = new FactorOverlap();
  int least = 0, greatest = 0;
  auto ok = [&](cudaError_t e, const char *what) {
    if (e == cudaSuccess) return true;
    set_error("factor_overlap_create: %s -> %s", what, cudaGetErrorString(e));
    factor_overlap_destroy(ov);          // releases whatever exists so far
    return false;
  };
  if (!ok(cudaDeviceGetStreamPriorityRange(&least, &greatest), "cudaDeviceGetStreamPriorityRange")) return -1;
  if (!ok(cudaStreamCreateWithPriority(&ov->main, cudaStreamNonBlocking, greatest), "cudaStreamCreateWithPriority")) return -1;
  for (int d = 0; d < FactorOverlap::MAX_DEPTH; ++d) {
    if (!ok(cudaStreamCreateWithPriority(&ov->side[d], cudaStreamNonBlocking, least), "cudaStreamCreateWithPriority")) return -1;
    if (!ok(cudaEventCreateWithFlags(&ov->fork[d], cudaEventDisableTiming), "cudaEventCreateWithFlags")) return -1;
    if (!ok(cudaEventCreateWithFlags(&ov->join[d], cudaEventDisableTiming), "cudaEventCreateWithFlags")) return -1;
  }
  if (!ok(cudaEventCreateWithFlags(&ov->enter, cudaEventDisableTiming), "cudaEventCreateWithFlags")) return -1;
  if (!ok(cudaEventCreateWithFlags(&ov->leave, cudaEventDisableTiming), "cudaEventCreateWithFlags")) return -1;
  *out = ov;
  return 0;
}

void factor_overlap_destroy(FactorOverlap *ov) {
  if (!ov) return;
  for (int d = 0; d < FactorOverlap::MAX_DEPTH; ++d) {
    if (ov->side[d]) cudaStreamDestroy(ov->side[d]);
    if (ov->fork[d]) cudaEventDestroy(ov->fork[d]);
    if (ov->join[d]) cudaEventDestroy(ov->join[d]);
  }
  if (ov->main) cudaStreamDestroy(ov->main);
  if (ov->enter) cudaEventDestroy(ov->enter);
  if (ov->leave) cudaEventDestroy(ov->leave);
  delete ov;
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
  ozaki_invalidate();
  return trtri_rec(f, 0, f.np);
}

int factor_potrf_inv(Factor &f) {
  GPB_REQUIRE(f.np % TILE == 0 && f.np >= TILE, "factor: padded size must be a multiple of 128");
  GPB_CUDA(cudaMemsetAsync(f.info, 0, sizeof(int), f.stream));
  f.l_pending = true;
  f.l_from = 0;
  ozaki_invalidate();
  return cholinv(f, 0, f.np, 0);
}

// Append: the leading h x h part of A / Mi holds a valid factor and inverse factor, the block rows from h on hold fresh rows of
// Ky (lower blocks).  Factorises them against the existing part.  h multiple of 128, 0 < h < np.
int factor_append(Factor &f, int h) {
  GPB_REQUIRE(f.np % TILE == 0 && h % TILE == 0 && h > 0 && h < f.np, "factor_append: bad split %d of %d", h, f.np);
  GPB_REQUIRE(!f.l_pending, "factor_append: finalize L first");
  GPB_CUDA(cudaMemsetAsync(f.info, 0, sizeof(int), f.stream));
  f.l_pending = true;
  f.l_from = h / TILE;
  ozaki_invalidate();
  return cholinv(f, 0, f.np, 0, h);
}

// Ky^-1 = M^T M: W[i][j] = sum_{k >= max(i,j)} M[k][i] M[k][j]; lower tiles only (diagonal tiles complete).
int factor_potri(Factor &f) {
  GPB_TRY(factor_finalize_L(f));   // W still holds the off-diagonal blocks of L
  GemmArgs g{f.Mi, f.np, f.Mi, f.np, f.W, f.np, f.np, f.np, f.np, 1.0, 0.0, 1, 1, 0};
  if (ozaki_min_n() > 0 && f.np >= ozaki_min_n()) return ozaki_gemm_launch(LAYOUT_COLK, LAYOUT_COLK, g, 2, 2, 0, f.stream);
  return gemm_launch(LAYOUT_COLK, LAYOUT_COLK, g, f.stream);
}

// Ky^-1 across an append at h (W11 = leading h x h block of W, lower tiles): with M = [M11 0; M21 M22],
//     Ky^-1 = M^T M = [M11^T M11 + M21^T M21   . ;  M22^T M21   M22^T M22]
// so the old inverse only needs the rank-r term of the new block rows added: O(N^2 r) instead of N^3 / 3.
// factor_potri_downdate removes the term of the block rows [h, np_old) of the OLD factor (the partially filled last block row,
// which the append recomputes) while those rows still exist; factor_potri_append adds the term of the new rows and fills the
// block rows from h on.
int factor_potri_downdate(Factor &f, int h, int np_old) {
  if (np_old <= h) return 0;
  const int ld = f.np;
  GemmArgs g{f.Mi + (size_t)h * ld, ld, f.Mi + (size_t)h * ld, ld, f.W, ld, h, h, np_old - h, -1.0, 1.0, 1, 0, 0};
  return gemm_launch(LAYOUT_COLK, LAYOUT_COLK, g, f.stream);
}

int factor_potri_append(Factor &f, int h) {
  GPB_REQUIRE(f.np % TILE == 0 && h % TILE == 0 && h > 0 && h < f.np, "factor_potri_append: bad split %d of %d", h, f.np);
  GPB_TRY(factor_finalize_L(f));   // W[h:, :] still holds L21 and the scratch of the trailing recursion
  const int ld = f.np, r = f.np - h;
  const double *M21 = f.Mi + (size_t)h * ld, *M22 = f.Mi + (size_t)h * ld + h;
  GemmArgs g;
  g = GemmArgs{M21, ld, M21, ld, f.W, ld, h, h, r, 1.0, 1.0, 1, 0, 0};                                  // W11 += M21^T M21
  GPB_TRY(gemm_launch(LAYOUT_COLK, LAYOUT_COLK, g, f.stream));
  g = GemmArgs{M22, ld, M21, ld, f.W + (size_t)h * ld, ld, r, h, r, 1.0, 0.0, 0, 1, 0};                 // W21 = M22^T M21
  GPB_TRY(gemm_launch(LAYOUT_COLK, LAYOUT_COLK, g, f.stream));
  g = GemmArgs{M22, ld, M22, ld, f.W + (size_t)h * ld + h, ld, r, r, r, 1.0, 0.0, 1, 1, 0};             // W22 = M22^T M22
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

// ---------------------------------------------------------------------------------------------------------------------
// Skinny triangular products for a handful of right-hand sides (the M = 1 .. 8 candidate calls that L-BFGS-B makes from
// every anchor point, optimizer.py:46-51): one pass over M serves all C vectors, so the cost is the HBM read of the
// triangle (8 N^2 / 2 bytes) instead of a 128-row padded GEMM.
//   Z[c][i] = sum_{k <= i} M[i][k] B[c][k]          (trmm_lower_skinny:   Vt = KxT M^T, row c = M k*_c)
//   U[c][j] = sum_{i >= j} M[i][j] Z[c][i]          (trmm_lower_t_skinny: Ut = Vt M,    row c = M^T (M k*_c))
// Fixed reduction orders -> bitwise reproducible.
// ---------------------------------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(256) trmm_lower_skinny_kernel(const double *__restrict__ M, int ld, int np,
                                                                const double *__restrict__ B, int ldb, double *__restrict__ Z,
                                                                int ldz) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= np) return;
  const int i = np - 1 - warp;             // long rows first
  const int kend = (i / TILE + 1) * TILE;  // the diagonal leaf block holds explicit zeros above the diagonal
  const double *row = M + (size_t)i * ld;
  double acc[C];
#pragma unroll
  for (int c = 0; c < C; ++c) acc[c] = 0.0;
  int k = lane;
  for (; k + 96 < kend; k += 128) {
    const double m0 = row[k], m1 = row[k + 32], m2 = row[k + 64], m3 = row[k + 96];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const double *b = B + (size_t)c * ldb + k;
      acc[c] = fma(m0, b[0], acc[c]);
      acc[c] = fma(m1, b[32], acc[c]);
      acc[c] = fma(m2, b[64], acc[c]);
      acc[c] = fma(m3, b[96], acc[c]);
    }
  }
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const double v = warp_sum(acc[c]);
    if (lane == 0) Z[(size_t)c * ldz + i] = v;
  }
}

constexpr int TRMM_T_SPLITS = 16;

// Column block cb (128 columns) has nb - cb row chunks below and on the diagonal; they are cut into runs of `run` consecutive
// chunks, one CTA each, so that every CTA streams about the same number of bytes (a fixed number of splits per column block
// left the CTAs of the short columns idle for half of the kernel).  run = ceil(nb / TRMM_T_SPLITS) keeps splits <= 16.
//   part[(c * S + s) * np + j] = sum over the rows of run s of column block j / 128 of  M[i][j] Z[c][i]
template <int C>
__global__ void __launch_bounds__(TILE) trmm_lower_t_skinny_partial_kernel(const double *__restrict__ M, int ld,
                                                                           const double *__restrict__ Z, int ldz,
                                                                           double *__restrict__ part, int np, int run) {
  const int cb = blockIdx.x, s = blockIdx.y, nb = np / TILE;
  const int rc0 = cb + s * run, rc1 = min(nb, rc0 + run);
  if (rc0 >= nb) return;
  const int j = cb * TILE + threadIdx.x;
  double acc[C];
#pragma unroll
  for (int c = 0; c < C; ++c) acc[c] = 0.0;
  const double *p = M + (size_t)(rc0 * TILE) * ld + j;
  const double *zz = Z + rc0 * TILE;
  const int rows = (rc1 - rc0) * TILE;
#pragma unroll 16
  for (int i = 0; i < rows; ++i) {
    const double m = p[(size_t)i * ld];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = fma(m, zz[(size_t)c * ldz + i], acc[c]);
  }
#pragma unroll
  for (int c = 0; c < C; ++c) part[((size_t)c * TRMM_T_SPLITS + s) * np + j] = acc[c];
}

__global__ void trmm_lower_t_skinny_reduce_kernel(const double *__restrict__ part, int np, int C, double *__restrict__ U, int ldu, int run) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y;
  if (j >= np || c >= C) return;
  const int nb = np / TILE, cb = j / TILE;
  const int splits = (nb - cb + run - 1) / run;   // runs that exist for this column block
  double acc = 0.0;
  for (int s = 0; s < splits; ++s) acc += part[((size_t)c * TRMM_T_SPLITS + s) * np + j];
  U[(size_t)c * ldu + j] = acc;
}

// C in {1, 2, 4, 8} right-hand sides stored as rows of B (row stride ldb); rows beyond the caller's count must be readable
// (zero padded).  want_u = 0: only Z.  part: >= 8 * TRMM_T_SPLITS * np doubles.
int factor_skinny_products(Factor &f, int c, const double *B, int ldb, double *Z, int ldz, double *U, int ldu, double *part) {
  GPB_REQUIRE(c == 1 || c == 2 || c == 4 || c == 8, "skinny products: C must be 1, 2, 4 or 8 (got %d)", c);
  const int np = f.np, nb = np / TILE;
  const int blocks = (np * 32 + 255) / 256;
  const int run = (nb + TRMM_T_SPLITS - 1) / TRMM_T_SPLITS;
  const dim3 gp(nb, TRMM_T_SPLITS), gr((np + 255) / 256, c);
#define GPB_SK(C_)                                                                                     \
  do {                                                                                                 \
    trmm_lower_skinny_kernel<C_><<<blocks, 256, 0, f.stream>>>(f.Mi, np, np, B, ldb, Z, ldz);           \
    GPB_CHECK_LAUNCH();                                                                                \
    count_launch();                                                                                    \
    if (U) {                                                                                           \
      trmm_lower_t_skinny_partial_kernel<C_><<<gp, TILE, 0, f.stream>>>(f.Mi, np, Z, ldz, part, np, run); \
      GPB_CHECK_LAUNCH();                                                                              \
      trmm_lower_t_skinny_reduce_kernel<<<gr, 256, 0, f.stream>>>(part, np, C_, U, ldu, run);          \
      GPB_CHECK_LAUNCH();                                                                              \
      count_launch(2);                                                                                 \
    }                                                                                                  \
  } while (0)
  if (c == 1) GPB_SK(1);
  else if (c == 2) GPB_SK(2);
  else if (c == 4) GPB_SK(4);
  else GPB_SK(8);
#undef GPB_SK
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
