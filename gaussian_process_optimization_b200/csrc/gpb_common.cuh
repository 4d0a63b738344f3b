// Shared definitions for libgpb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <mutex>

#include "../../include/gpb200.h"

namespace gpb {

constexpr int TILE = 128;  // everything N x N is padded to a multiple of this; also the Cholesky leaf size
constexpr double SQRT5 = 2.23606797749978969640917366873128;

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

// ---- error plumbing -------------------------------------------------------------------------------------------------
void set_error(const char *fmt, ...);
extern std::atomic<long long> g_launches;  // kernels launched by this library (bench.py's gpu_launches); models may be driven from several host threads
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define GPB_CUDA(call)                                                                         \
  do {                                                                                         \
    cudaError_t e_ = (call);                                                                   \
    if (e_ != cudaSuccess) {                                                                   \
      gpb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));    \
      return -1;                                                                               \
    }                                                                                          \
  } while (0)

#define GPB_CHECK_LAUNCH()                                                                     \
  do {                                                                                         \
    cudaError_t e_ = cudaGetLastError();                                                       \
    if (e_ != cudaSuccess) {                                                                   \
      gpb::set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e_)); \
      return -1;                                                                               \
    }                                                                                          \
  } while (0)

#define GPB_REQUIRE(cond, ...)            \
  do {                                    \
    if (!(cond)) {                        \
      gpb::set_error(__VA_ARGS__);        \
      return -2;                          \
    }                                     \
  } while (0)

#define GPB_TRY(call)        \
  do {                       \
    int r_ = (call);         \
    if (r_ != 0) return r_;  \
  } while (0)

// Opt-in dynamic shared memory must be configured once per kernel AND per device (function attributes are per context):
// `done` is a per-call-site bit mask over device ordinals.  Models may be driven from several host threads (concurrent restarts),
// so the first use is serialised: the bit is published (release) only after the configuring thread has finished, and a thread
// that finds it unset waits on the mutex instead of launching an unconfigured kernel.
//   static FuncConfigMask configured;  FuncConfigOnce once(configured);  if (once.needed) { cudaFuncSetAttribute(...); }
typedef std::atomic<unsigned long long> FuncConfigMask;
std::mutex &func_config_mutex();
struct FuncConfigOnce {
  FuncConfigMask &done;
  unsigned long long bit = 0;
  bool needed = true;
  std::unique_lock<std::mutex> lock;
  explicit FuncConfigOnce(FuncConfigMask &d) : done(d) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return;     // unknown device: configure every time
    bit = 1ull << dev;
    if (done.load(std::memory_order_acquire) & bit) {
      needed = false;
      return;
    }
    lock = std::unique_lock<std::mutex>(func_config_mutex());
    needed = !(done.load(std::memory_order_acquire) & bit);
  }
  ~FuncConfigOnce() {
    if (needed && bit) done.fetch_or(bit, std::memory_order_release);
  }
};

// ---- programmatic dependent launch (PDL) -----------------------------------------------------------------------------------
// The factorisation is a chain of several hundred short dependent kernels on one stream.  Kernels launched through launch_pdl
// may be scheduled while the previous kernel of the stream drains (its CTAs signal pdl_trigger() at their start; the hardware
// acts once every CTA of that grid has signalled or exited) and run their prologue (index arithmetic, shared-memory set-up) in
// that shadow; pdl_wait() then blocks until the previous grid has completed and its memory is visible, so it must precede the
// first global access.  In a kernel launched the ordinary way both instructions are no-ops.  GPB_PDL=0 disables the attribute.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// ---- stationary covariance functions of the scaled squared distance ---------------------------------------------------
// k(r) and k'(r)/r for r = sqrt(r2).  The reference forms dK_dr * inv_dist (stationary.py:227-230,251-258) with inv_dist := 0
// where r == 0; k'(r)/r is finite at 0 and is only ever multiplied by (x - x') which vanishes there, so the closed form
// gives the same sums without the division.
//   RBF      (rbf.py:50-54):            k = v exp(-r^2/2),                        k'/r = -k
//   Matern52 (stationary.py:575-579):   k = v (1 + s5 r + 5/3 r^2) exp(-s5 r),    k'/r = -(5/3) v (1 + s5 r) exp(-s5 r)
// r2 can reach 1e300+ when an optimiser step drives a lengthscale towards 0 (scaled coordinates are clamped at 1e150 in
// scale_transpose_kernel); the covariance is exactly 0 there, but (1 + s + ..) * exp(-s) would evaluate inf * 0.  Clamping r2
// keeps every factor finite (NaN stays NaN: the comparison is false for it, and a NaN kernel matrix ends in the jitchol
// LinAlgError like in the reference).
__device__ __forceinline__ double clamp_r2(double r2) { return r2 > 1e300 ? 1e300 : r2; }

template <int KIND>
__device__ __forceinline__ double cov_k(double r2, double variance) {
  r2 = clamp_r2(r2);
  if (KIND == GPB_KERN_RBF) {
    return variance * exp(-0.5 * r2);
  } else {
    const double r = sqrt(r2);
    const double s = SQRT5 * r;
    return variance * (1.0 + s + (5.0 / 3.0) * r2) * exp(-s);
  }
}

template <int KIND>
__device__ __forceinline__ void cov_k_dk(double r2, double variance, double &k, double &dk_over_r) {
  r2 = clamp_r2(r2);
  if (KIND == GPB_KERN_RBF) {
    k = variance * exp(-0.5 * r2);
    dk_over_r = -k;
  } else {
    const double r = sqrt(r2);
    const double s = SQRT5 * r;
    const double e = variance * exp(-s);
    k = (1.0 + s + (5.0 / 3.0) * r2) * e;
    dk_over_r = -(5.0 / 3.0) * (1.0 + s) * e;
  }
}

// ---- Gower product kernel (stationary.py:116-135): accumulate over dimensions, then evaluate once ---------------------------
// RBF:      s += r^2                       value = vpow exp(-s / 2)
// Matern52: s += r, p *= 1 + s5 r + 5/3 r^2  value = vpow p exp(-s5 s)
template <int KIND>
__device__ __forceinline__ void gower_accumulate(double r, double &s, double &p) {
  if (KIND == GPB_KERN_RBF) {
    s = fma(r, r, s);
  } else {
    s += r;
    p *= fma(r, fma(5.0 / 3.0, r, SQRT5), 1.0);
  }
}
template <int KIND>
__device__ __forceinline__ double gower_value(double s, double p, double vpow) {
  if (KIND == GPB_KERN_RBF) return vpow * exp(-0.5 * s);
  return vpow * p * exp(-SQRT5 * s);
}

// ---- reductions -----------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum over a block of NT threads (NT multiple of 32, <= 1024); result valid in thread 0.  `scratch` holds >= 32 doubles.
// Fixed shuffle/tree order -> bitwise reproducible run to run.
template <int NT>
__device__ __forceinline__ double block_sum(double v, double *scratch) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) scratch[w] = v;
  __syncthreads();
  double r = 0.0;
  if (w == 0) {
    r = (l < NT / 32) ? scratch[l] : 0.0;
    r = warp_sum(r);
  }
  return r;
}

// ---- factorisation state shared by the linalg drivers ----------------------------------------------------------------
// All three matrices are Np x Np row-major with leading dimension Np (Np = round_up(N, TILE)); rows/cols >= N are the
// identity so that padded problems factor to [L 0; 0 I].
// Streams and events of the two-stream factorisation schedule (gpb_chol.cu: cholinv).
struct FactorOverlap {
  static constexpr int MAX_DEPTH = 10;
  cudaStream_t main = 0;               // high priority: the critical path of a fit runs here
  cudaStream_t side[MAX_DEPTH] = {};   // low priority, one per recursion depth
  cudaEvent_t fork[MAX_DEPTH] = {}, join[MAX_DEPTH] = {};
  cudaEvent_t enter = nullptr, leave = nullptr;  // hand-over between the caller's stream and `main`
  int min_n = 512;                     // smallest block size whose T21 product is forked
};
int factor_overlap_create(FactorOverlap **out);
void factor_overlap_destroy(FactorOverlap *ov);

struct Factor {
  int n = 0;    // logical size
  int np = 0;   // padded size
  double *A = nullptr;  // in: Ky (lower read); out: L in the lower triangle (diagonal leaf blocks have a zeroed upper part)
  double *Mi = nullptr; // out: L^-1 (lower; diagonal leaf blocks explicit zeros above the diagonal)
  double *W = nullptr;  // scratch during potrf; out of potri: Ky^-1 (lower tiles valid, diagonal tiles full)
  int *info = nullptr;  // device int: 0 or (1 + index of first non-positive pivot)
  double *part = nullptr;  // scratch for GEMV partials: (np / TILE) * np doubles
  cudaStream_t stream = 0;
  FactorOverlap *ov = nullptr;  // non-null: two-stream schedule (stream must then be ov->main)
  bool l_pending = false;       // the off-diagonal blocks of L still sit in W (factor_finalize_L moves them into A)
  int l_from = 0;               // ... for the block rows >= l_from (0 after a full factorisation, h / 128 after an append)
};

// gemm engine (gpb_gemm.cu)
enum { LAYOUT_ROWK = 0, LAYOUT_COLK = 1 };
struct GemmArgs {
  const double *A; int lda;
  const double *B; int ldb;
  double *C; int ldc;
  int M, N, K;
  double alpha, beta;
  int tri_out;   // 1: only tiles with tile_col <= tile_row are computed
  int klo_mode;  // 0: 0       1: row0        2: col0
  int khi_mode;  // 0: K       1: row0 + 128  2: col0 + 128
  int band = 0;  // tri_out: block rows per band of the tile enumeration (0 = default 12; GPB_TRI_BAND)
};
int gemm_launch(int layout_a, int layout_b, const GemmArgs &g, cudaStream_t s);
int gemm_profile_enable(int on);
bool gemm_profile_is_on();
int gemm_force_config(int cfg);  // 0 auto, 1 BIG (64x128), 2 MID (64x64), 3 SMALL (32x32)
int gemm_profile_collect(double *ms, double *flops, long long *launches);
int gemm_profile_last(double *ms, double *flops);

// int8 tensor-core (tcgen05) Ozaki engine for the large products (gpb_ozaki.cu); experimental, off unless configured
int ozaki_gemm_launch(int layout_a, int layout_b, const GemmArgs &g, int tri_a, int tri_b, int slices, cudaStream_t s, int cache_b = 0);
void ozaki_invalidate();                     // a factorisation ran: cached digit planes of L^-1 are stale
int ozaki_min_n();                           // products of the recursion with n >= this go through the engine (0 = off)
int ozaki_configure(int min_n, int slices);
void ozaki_release_stream(cudaStream_t st);   // the stream is about to be destroyed: free the engine's plane workspace keyed by it
void ozaki_suppress(int on);                 // thread-local: 1 = the calling thread's products stay on the DMMA engine until ozaki_suppress(0)
int ozaki_predict_planes();                 // planes of the predictive products: 8 digits, or 18 moduli in modular mode
int ozaki_crt_bits(int nmod, long long k);   // modular mode (gpb_crt.cuh): bits per operand; host restatements for the CPU tests
int ozaki_crt_host_residues(const double *A, int rows, int k, int nmod, int beta, signed char *planes, double *scale);
int ozaki_crt_host_combine(const int *sums, long long count, int nmod, double *X);

// linalg drivers (gpb_chol.cu)
int factor_potrf_inv(Factor &f);                 // A -> L, Mi = L^-1, *info
int factor_trtri(Factor &f);                     // A holds a lower-triangular L (diag blocks clean) -> Mi = L^-1
int factor_potri(Factor &f);                     // W = Mi^T Mi (lower tiles)
int factor_finalize_L(Factor &f);                // off-diagonal blocks of L: W -> A (idempotent)
int factor_append(Factor &f, int h);             // leading h x h part already factorised: factor the block rows from h on
// N <= 128, D <= 32, one output: Ky build + factor + inverse + solve + log det (+ Ky^-1 and the gradient sums) in one kernel
int launch_tiny_fit(int kind, const double *X, const double *ls_host, double *XsT, double *ls_dev, double *inv_ls_dev, int n, int d,
                    double variance, double diag_add, const double *y, Factor &f, double *z, double *alpha, double *scal, int want_grad);
int factor_potri_downdate(Factor &f, int h, int np_old);  // W11 -= (rows [h, np_old) of Mi)^T (same rows): before an append
int factor_potri_append(Factor &f, int h);       // W: inverse of the leading h block (downdated) -> Ky^-1 of the extended matrix
int factor_solve(Factor &f, const double *Y, int p, double *z, double *alpha);  // alpha = Mi^T (Mi Y);  Y, z, alpha: np x p col-major (p vectors of np)
int factor_logdet(Factor &f, double *out_dev);   // 2 sum log L_ii, i < n
// Z[c] = Mi B[c], U[c] = Mi^T Z[c] for c in {1,2,4,8} vectors stored as rows (U may be NULL); part: 8 * 16 * np doubles
int factor_skinny_products(Factor &f, int c, const double *B, int ldb, double *Z, int ldz, double *U, int ldu, double *part);
int launch_copy2d(double *dst, int ldd, const double *src, int lds, int rows, int cols, cudaStream_t s);
int launch_symmetrize_lower(double *A, int ld, int n, cudaStream_t s);

}  // namespace gpb
