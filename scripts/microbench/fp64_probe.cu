// Micro-probes for the fp64 scalar pipe of one SM (B200): throughput vs warps, dependent latency, barrier, division.
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dfma_tp(double *out, int iters, long long *cyc) {
  double a[ILP];
  const double x = 1.0000001, y = 1e-9 * threadIdx.x;
#pragma unroll
  for (int i = 0; i < ILP; ++i) a[i] = i + threadIdx.x;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = fma(a[i], x, y);
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void barrier_cost(long long *cyc, int iters) {
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void div_lat(double *out, int iters, long long *cyc) {
  double v = 3.0 + threadIdx.x;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) v = 1.0 / v + 2.0;
  const long long t1 = clock64();
  out[threadIdx.x] = v;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void lds_bcast_chain(double *out, int iters, long long *cyc) {
  __shared__ double buf[256];
  buf[threadIdx.x] = threadIdx.x;
  __syncthreads();
  double v = 0;
  int idx = threadIdx.x & 127;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    v += buf[idx];
    idx = (idx + (int)v) & 127;
  }
  const long long t1 = clock64();
  out[threadIdx.x] = v;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

// publish -> barrier -> load -> rcp -> mul -> fma -> publish : the leaf's per-column chain
__global__ void chain(double *out, int iters, long long *cyc) {
  __shared__ double buf[2][128];
  double s = 2.0 + threadIdx.x * 1e-3;
  if (threadIdx.x < 128) buf[0][threadIdx.x] = s;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const double *cb = buf[it & 1];
    const double piv = cb[it & 127];
    const double r = -1.0 / piv;
    const double w = cb[threadIdx.x & 127] * r;
    s = fma(w, cb[(threadIdx.x + 1) & 127], s) + 3.0;
    if (threadIdx.x < 128) buf[(it + 1) & 1][threadIdx.x] = s;
    __syncthreads();
  }
  const long long t1 = clock64();
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
  double *out;
  long long *cyc, h[64];
  cudaMalloc(&out, 1 << 20);
  cudaMalloc(&cyc, 64 * 8);
  const int iters = 2000;
  for (int warps : {1, 2, 4, 8, 16, 32}) {
    dfma_tp<8><<<1, warps * 32>>>(out, iters, cyc);
    cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("dfma ILP8  warps/SM=%2d: %.2f cycles per warp-DFMA per scheduler-slot, %.1f FMA/clk/SM\n", warps,
           (double)h[0] / (iters * 8.0), warps * 32.0 * iters * 8 / h[0]);
  }
  dfma_tp<1><<<1, 32>>>(out, iters, cyc);
  cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("dfma dependent latency: %.2f cycles\n", (double)h[0] / iters);
  dfma_tp<2><<<1, 32>>>(out, iters, cyc);
  cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("dfma ILP2 one warp: %.2f cycles per DFMA\n", (double)h[0] / (iters * 2));
  dfma_tp<4><<<1, 32>>>(out, iters, cyc);
  cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("dfma ILP4 one warp: %.2f cycles per DFMA\n", (double)h[0] / (iters * 4));
  dfma_tp<16><<<1, 32>>>(out, iters, cyc);
  cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("dfma ILP16 one warp: %.2f cycles per DFMA\n", (double)h[0] / (iters * 16));
  dfma_tp<16><<<1, 256>>>(out, iters, cyc);
  cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("dfma ILP16 8 warps: %.2f cycles per DFMA per warp, %.1f FMA/clk/SM\n", (double)h[0] / (iters * 16), 256.0 * iters * 16 / h[0]);
  for (int threads : {64, 128, 256, 512}) {
    barrier_cost<<<1, threads>>>(cyc, iters);
    cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("__syncthreads with %d threads: %.1f cycles\n", threads, (double)h[0] / iters);
  }
  div_lat<<<1, 32>>>(out, iters, cyc);
  cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("1.0/v + 2.0 dependent chain: %.1f cycles per iteration\n", (double)h[0] / iters);
  lds_bcast_chain<<<1, 32>>>(out, iters, cyc);
  cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("LDS dependent chain (+DADD, F2I): %.1f cycles per iteration\n", (double)h[0] / iters);
  chain<<<1, 256>>>(out, iters, cyc);
  cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("leaf-like column chain (LDS, rcp, mul, fma, STS, barrier), 256 threads: %.1f cycles per column\n", (double)h[0] / iters);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
