// Do the fp64 FMA lanes (DFMA) and the fp64 tensor path (DMMA.8x8x4) of one B200 SM run concurrently?
// Each warp runs either an independent-accumulator DMMA loop or a DFMA loop; we report FMA / clk / SM for pure and mixed mixes.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// mode bit per warp: dmma_warps = number of warps (lowest ids) doing DMMA, the rest DFMA
__global__ void mix(double *out, int iters, int dmma_warps, long long *cyc) {
  const int warp = threadIdx.x >> 5;
  double acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = i + threadIdx.x * 1e-3;
  const double a = 1.0000001 + threadIdx.x * 1e-9, b = 0.9999999;
  __syncthreads();
  const long long t0 = clock64();
  if (warp < dmma_warps) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) dmma884(acc[2 * i], acc[2 * i + 1], a, b);   // 8 independent DMMAs = 8 * 256 FMA per warp
    }
  } else {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);                 // 64 DFMA = 64 * 32 FMA per warp
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  const long long t2 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if ((threadIdx.x & 31) == 0) cyc[warp] = t1 - t0;
  if (threadIdx.x == 0) cyc[63] = t2 - t0;
}

int main() {
  double *out;
  long long *cyc, h[64];
  cudaMalloc(&out, 1 << 20);
  cudaMalloc(&cyc, 64 * 8);
  const int iters = 4000;
  const int mixes[][2] = {{8, 0}, {0, 8}, {4, 4}, {16, 0}, {0, 16}, {8, 8}, {4, 12}, {12, 4}, {4, 8}, {4, 16}};
  for (auto &mx : mixes) {
    const int nd = mx[0], nf = mx[1];
    mix<<<1, (nd + nf) * 32>>>(out, iters, nd, cyc);
    cudaMemcpy(h, cyc, 64 * 8, cudaMemcpyDeviceToHost);
    const double total = (double)h[63];
    const double fma_dmma = (double)nd * iters * 8 * 256, fma_dfma = (double)nf * iters * 64 * 32;
    double td = 0, tf = 0;
    for (int w = 0; w < nd; ++w) td = h[w] > td ? h[w] : td;
    for (int w = nd; w < nd + nf; ++w) tf = h[w] > tf ? h[w] : tf;
    printf("DMMA warps %2d, DFMA warps %2d: total %.0f cycles -> %.1f FMA/clk/SM overall", nd, nf, total, (fma_dmma + fma_dfma) / total);
    if (nd) printf("; DMMA part %.1f FMA/clk (%.0f cyc)", fma_dmma / td, td);
    if (nf) printf("; DFMA part %.1f FMA/clk (%.0f cyc)", fma_dfma / tf, tf);
    printf("\n");
  }
  printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
