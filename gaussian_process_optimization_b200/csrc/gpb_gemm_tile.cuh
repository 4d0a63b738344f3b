// Operand staging and fragment access of the fp64 DMMA GEMM engine, shared by the stand-alone kernel (gpb_gemm.cu) and the
// cooperative diagonal-block solver (gpb_chol.cu), which must reproduce the engine's arithmetic tile for tile.
#pragma once
#include "gpb_common.cuh"

namespace gpb {

constexpr int BK_MIN = 16;  // every k-range is a multiple of this (and of every BK used below)

__device__ __forceinline__ void cp_async16(double *smem_dst, const double *gmem_src) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// Stage one ROWS(row) x 16(k) operand tile.  base points at element (row0, 0) [ROWK] or (0, row0) [COLK] of the operand,
// kk is the k offset of this tile.
template <int LAYOUT, int ROWS, int THREADS, int BK>
__device__ __forceinline__ void load_tile(double *s, const double *__restrict__ base, int ld, int kk, int tid) {
  constexpr int CHUNKS = ROWS * BK / 2;  // 16-byte chunks
  constexpr int LD_ROWK = BK + 4;        // operand stored [row][k]  (k contiguous)
  static_assert(CHUNKS % THREADS == 0, "tile/threads mismatch");
  if (LAYOUT == LAYOUT_ROWK) {
    constexpr int CPR = BK / 2;  // chunks per row
#pragma unroll
    for (int i = 0; i < CHUNKS / THREADS; ++i) {
      const int chunk = tid + i * THREADS;
      const int row = chunk / CPR, c = chunk % CPR;
      cp_async16(s + row * LD_ROWK + 2 * c, base + (size_t)row * ld + kk + 2 * c);
    }
  } else {
    constexpr int LD = ROWS + 4, CPR = ROWS / 2;  // chunks per k-row
#pragma unroll
    for (int i = 0; i < CHUNKS / THREADS; ++i) {
      const int chunk = tid + i * THREADS;
      const int kr = chunk / CPR, c = chunk - kr * CPR;
      cp_async16(s + kr * LD + 2 * c, base + (size_t)(kk + kr) * ld + 2 * c);
    }
  }
}

template <int LAYOUT, int ROWS, int BK>
__device__ __forceinline__ double frag(const double *s, int row, int k) {
  return (LAYOUT == LAYOUT_ROWK) ? s[row * (BK + 4) + k] : s[k * (ROWS + 4) + row];
}

template <int LAYOUT, int ROWS, int BK>
__host__ __device__ constexpr int tile_doubles() {
  return (LAYOUT == LAYOUT_ROWK) ? ROWS * (BK + 4) : BK * (ROWS + 4);
}


}  // namespace gpb
