// Leaf micro-benchmark: times the 128x128 factor+inverse leaves alone and (with -DGPB_LEAF_TIMING) prints the cycles of each
// phase of the blocked leaf.  Build (from the repo root):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DGPB_LEAF_TIMING -o scripts/microbench/leaf_bench \
//        scripts/microbench/leaf_bench.cu gaussian_process_optimization_b200/csrc/gpb_{api,gemm,kernels,predict}.cu
#include "../../gaussian_process_optimization_b200/csrc/gpb_chol.cu"

#include <vector>

using namespace gpb;

int main() {
  const int n = TILE;
  std::vector<double> hA(n * n), hL(n * n), hM(n * n);
  // SPD test block: RBF-like kernel matrix + noise
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      const double d = (i - j) / 9.0;
      hA[i * n + j] = exp(-0.5 * d * d) + (i == j ? 1e-2 : 0.0);
    }
  double *A, *Mi, *W;
  int *info;
  cudaMalloc(&A, n * n * 8);
  cudaMalloc(&Mi, n * n * 8);
  cudaMalloc(&W, n * n * 8);
  cudaMalloc(&info, 4);
  Factor f;
  f.n = n; f.np = n; f.A = A; f.Mi = Mi; f.W = W; f.info = info;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int variant = 0; variant < 2; ++variant) {
    g_leaf_variant = variant;
    float best = 1e9;
    for (int rep = 0; rep < 6; ++rep) {
      cudaMemcpy(A, hA.data(), n * n * 8, cudaMemcpyHostToDevice);
      cudaMemset(info, 0, 4);
      cudaDeviceSynchronize();
      cudaEventRecord(e0);
      if (launch_leaf(f, 0, 0) != 0) { printf("launch failed: %s\n", gpb_last_error()); return 1; }
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) best = ms;
    }
    cudaMemcpy(hL.data(), A, n * n * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(hM.data(), Mi, n * n * 8, cudaMemcpyDeviceToHost);
    int hinfo;
    cudaMemcpy(&hinfo, info, 4, cudaMemcpyDeviceToHost);
    // residuals: |L L^T - A|, |M L - I|
    double r1 = 0, r2 = 0;
    for (int i = 0; i < n; ++i)
      for (int j = 0; j <= i; ++j) {
        double s = 0, t = 0;
        for (int k = 0; k < n; ++k) { s += hL[i * n + k] * hL[j * n + k]; t += hM[i * n + k] * hL[k * n + j]; }
        r1 = fmax(r1, fabs(s - hA[i * n + j]));
        r2 = fmax(r2, fabs(t - (i == j ? 1.0 : 0.0)));
      }
    printf("variant %d: %.2f us  info %d  |LL^T-A| %.2e  |ML-I| %.2e  (%s)\n", variant, best * 1e3, hinfo, r1, r2, cudaGetErrorString(cudaGetLastError()));
#ifdef GPB_LEAF_TIMING
    if (variant == 1) {
      long long c[64];
      cudaMemcpyFromSymbol(c, g_leaf_clk, sizeof(c));
      printf("load %lld\n", c[1] - c[0]);
      for (int k = -1; k < 8; ++k) {
        const long long *q = c + 2 + 4 * (k + 1);
        printf("k=%d  panel %lld  warp 0 (update + diag k+1 | inverse row share) %lld  waiting for warps 1-7 %lld\n", k, q[1] - q[0],
               q[2] - q[1], q[3] - q[2]);
      }
      printf("total %lld cycles\n", c[41] - c[0]);
    }
#endif
  }
  return 0;
}
