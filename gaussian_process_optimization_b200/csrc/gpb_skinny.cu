// The M = 1 .. 8 predictive call as ONE persistent cooperative kernel.
//
// What it serves: the f_df calls bounded L-BFGS-B makes from every anchor point (GPyOpt/GPyOpt/optimization/optimizer.py:46-51 ->
// acquisition_function_withGradients -> GPModel.predict_withGradients, gpmodel.py:131-142 -> GP.predict + predictive_gradients,
// core/gp.py:297-354,407-454 -> PosteriorExact._raw_predict, posterior.py:273-302) -- hundreds of them per BO step, and (through
// the host's LockstepEvaluator) up to 8 candidates per call.
//
// Such a call is bound by HBM: Z = M k* and U = M^T Z (M = L^-1, lower triangular) each stream the triangle once, 2 x 8 N^2 / 2
// bytes, whatever the number of candidates up to 8.  The multi-kernel route (scale_transpose, kmat, trmm_lower_skinny, its
// transposed partial + reduce, skinny moments partial + reduce: 7 dependent launches around two streaming kernels) leaves a third of
// the call to launch gaps and small kernels.  Here one grid of co-resident CTAs (cooperative launch, grid-wide barriers) does
//   P0  k*_c[j] = k(x*_c, x_j) and k'/r, all training points j            (posterior.py:275; rbf.py:50-54, stationary.py:575-579)
//   P1  Z_c = M k*_c           row panels of 32 rows x 128-column chunks    (the dtrtrs of posterior.py:293 as a product with L^-1)
//   P2  U_c = M^T Z_c          the same chunks walked column block by column block     (woodbury_inv product, core/gp.py:450-451)
//   P3  mu = k*^T alpha, var = base - |Z|^2, d mu / dx* and d var / dx*     (posterior.py:276,294-295; core/gp.py:431-434,450-453;
//                                                                            stationary.py:354-364)
// with the 32 KB chunks of the triangle dealt to the CTAs in contiguous, equal runs (+-1 chunk) and streamed through a 3-stage
// shared-memory ring by the TMA engine -- ONE cp.async.bulk.tensor.2d per chunk (tensor map over M, box 128 x 32 doubles) completing on
// an mbarrier -- every lane owning four columns of a chunk (two 16-byte pieces 512 bytes apart, so that a warp's shared loads are conflict free).  (History, ncu r2c / r2d: loading the chunks into registers kept too few
// bytes in flight, 4.8 of 6.5 TB/s; 32 one-row bulk copies of 1 KB per chunk made the TMA engine's per-operation cost the bound, 2.7.)
// All partial sums are combined in fixed orders that do not depend on the number of candidates sharing the call: candidate c's
// results are bit-identical whether it is evaluated alone or with seven others, and run to run.
#include <cooperative_groups.h>
#include <cuda.h>

#include <algorithm>

#include "gpb_common.cuh"
#include "gpb_kernels.cuh"

namespace cg = cooperative_groups;

namespace gpb {

constexpr int SK_THREADS = 256;
constexpr int SK_WARPS = SK_THREADS / 32;
constexpr int SK_PH = 32;        // rows of a panel (4 per warp)
constexpr int SK_SLOTS1 = 4;     // CTAs a row panel's chunks can be spread over     (G <= 4 (nb + 1))
constexpr int SK_SLOTS2 = 12;    // CTAs a column block's chunks can be spread over
constexpr int SK_QC = 8;         // input dimensions per pass of the gradient sums
constexpr int SK_STAGES = 3;     // shared-memory ring: stages of one chunk (32 rows x 128 doubles) + the C x 128 right-hand sides
constexpr int SK_MTILE = SK_PH * TILE;   // doubles of a chunk

__device__ __forceinline__ uint32_t sk_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sk_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void sk_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sk_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done)
                 : "r"(bar), "r"(parity)
                 : "memory");
  }
}
// one box of a 2-d tensor map (inner coordinate c0 = column, c1 = row) global -> shared through the TMA engine.  `policy`: an L2
// cache policy (createpolicy): the gigabyte of M streamed per pass is marked evict-first so that it does not flush the kilobytes
// everything else lives on (the right-hand sides, the partial sums, and the kernel's own code: after the first version's passes
// the cold instruction fetches of the reduction code alone cost 25 us, r2 phase stamps)
__device__ __forceinline__ void sk_tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(bar), "l"(policy)
      : "memory");
}
__device__ __forceinline__ uint64_t sk_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t sk_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void sk_fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
// the same for shared memory only (a stage of the ring read by the warps, then refilled by the TMA engine): no MEMBAR.GPU
__device__ __forceinline__ void sk_fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void sk_stamp(unsigned long long *dbg, int k) {
  if (dbg && blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    dbg[k] = t;
  }
}

struct SkinnyParams {
  const double *M;          // L^-1, np x np row-major (lower; diagonal 128-blocks carry explicit zeros above the diagonal)
  const double *XT;         // [d][np] training inputs divided by the lengthscales
  const double *Xc;         // [mc][d] raw candidates (device)
  const double *ls, *inv_ls, *alpha;
  double *Kx, *Dk, *Z, *U;  // [C][np] scratch: k(x*_c, x_j), k'(r) / r, Z_c = M k*_c and U_c = M^T Z_c once their partial sums are added
  double *part1;            // [np / 32][SK_SLOTS1][C][32]
  double *part2;            // [np / 128][SK_SLOTS2][C][128]
  double *part3;            // [G][C][2 + 2 d]
  unsigned int *ticket;
  unsigned long long *dbg;  // GPB_SKINNY_DBG=1: globaltimer stamps of CTA 0 at the phase boundaries (NULL otherwise)
  double *mu, *var, *dmu, *dvar;
  double variance, var_base;
  int np, n, d, mc, level;  // level 1: mean + variance; 2: + both gradients; 3: mean and its gradient only
};

// chunks are numbered 0 .. T - 1; CTA g of G owns [g T / G, (g + 1) T / G)
__device__ __forceinline__ int sk_owner(long long u, int G, long long T) { return (int)(((u + 1) * G - 1) / T); }
// first chunk of row panel p (32 rows; its chunks are the column blocks 0 .. p / 4)
__device__ __forceinline__ long long sk_prefix1(int p) {
  const long long bi = p >> 2, s = p & 3;
  return 2 * bi * (bi + 1) + s * (bi + 1);
}
// first chunk of column block cb (its chunks are the row panels 4 cb .. P - 1)
__device__ __forceinline__ long long sk_prefix2(int cb, int P) { return (long long)cb * P - 2ll * cb * (cb - 1); }

// Z_c[i] for the 32 rows of panel p = sum of the slots the owning CTAs wrote, in slot order.  All slot loads are issued before the
// first addition (clamped index, value masked afterwards): issued one by one behind their additions they serialise a memory
// round trip each, which is what the tail of this kernel is made of.
template <int C>
__device__ __forceinline__ double sk_zsum(const SkinnyParams &a, int c, int i, int G, long long T) {
  const int p = i >> 5, r = i & 31;
  const int ns = sk_owner(sk_prefix1(p + 1) - 1, G, T) - sk_owner(sk_prefix1(p), G, T);     // last slot index
  double v[SK_SLOTS1];
#pragma unroll
  for (int s = 0; s < SK_SLOTS1; ++s) v[s] = a.part1[(((size_t)p * SK_SLOTS1 + min(s, ns)) * C + c) * SK_PH + r];
  double z = 0.0;
#pragma unroll
  for (int s = 0; s < SK_SLOTS1; ++s) z += (s <= ns) ? v[s] : 0.0;
  return z;
}
template <int C>
__device__ __forceinline__ double sk_usum(const SkinnyParams &a, int c, int j, int G, long long T, int P) {
  const int cb = j >> 7, r = j & 127;
  const int ns = sk_owner(sk_prefix2(cb + 1, P) - 1, G, T) - sk_owner(sk_prefix2(cb, P), G, T);
  double v[SK_SLOTS2];
#pragma unroll
  for (int s = 0; s < SK_SLOTS2; ++s) v[s] = a.part2[(((size_t)cb * SK_SLOTS2 + min(s, ns)) * C + c) * TILE + r];
  double u = 0.0;
#pragma unroll
  for (int s = 0; s < SK_SLOTS2; ++s) u += (s <= ns) ? v[s] : 0.0;
  return u;
}

template <int KIND, int C>
__global__ void __launch_bounds__(SK_THREADS, 1) skinny_fused_kernel(const __grid_constant__ CUtensorMap mapM,
                                                                      const __grid_constant__ CUtensorMap mapK, const SkinnyParams a) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(128) unsigned char sk_dyn[];   // ring: SK_STAGES x (chunk of M | C x 128 right-hand sides)
  __shared__ double xcs[8 * 64];              // candidates divided by the lengthscales, [c][q]
  __shared__ double zs[2][C * SK_PH];         // Z of the current / next panel (P2)
  __shared__ __align__(8) unsigned long long full_bar[SK_STAGES];
  __shared__ int s_last;
  const int tid = threadIdx.x, lane = tid & 31;
  // the warp index through a shuffle: the compiler then KNOWS it is uniform across the warp and keeps the shuffle trees inside
  // warp-indexed loops on the converged fast path (without it every SHFL of the tail went through WARPSYNC.COLLECTIVE: 23 us, r2)
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int g = blockIdx.x, G = gridDim.x;
  const int np = a.np, n = a.n, d = a.d, nb = np / TILE, P = np / SK_PH;
  const long long T = 2ll * nb * (nb + 1);
  const bool want_var = a.level == 1 || a.level == 2, want_g = a.level >= 2, want_dvar = a.level == 2;
  constexpr int STAGE_DOUBLES = SK_MTILE + C * TILE;
  double *ring = reinterpret_cast<double *>((reinterpret_cast<uintptr_t>(sk_dyn) + 127) & ~uintptr_t(127));   // TMA destinations: 128-byte aligned
  double *wred = ring + (size_t)SK_STAGES * STAGE_DOUBLES;    // [SK_WARPS][C][128]: the warps' sums of a column block (P2)
  if (tid == 0) {
    for (int st = 0; st < SK_STAGES; ++st) sk_mbar_init(sk_smem_u32(&full_bar[st]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }

  sk_stamp(a.dbg, 0);
  // ---- P0: scaled candidates, then k* and k'/r over all training points -----------------------------------------------------
  for (int e = tid; e < C * d; e += SK_THREADS) {
    const int c = e / d, q = e - c * d;
    double v = (c < a.mc) ? a.Xc[(size_t)c * d + q] / a.ls[q] : 0.0;
    v = v > 1e150 ? 1e150 : (v < -1e150 ? -1e150 : v);        // as scale_transpose_kernel
    xcs[e] = v;
  }
  __syncthreads();
  for (int j = g * SK_THREADS + tid; j < np; j += G * SK_THREADS) {
    double r2[C];
#pragma unroll
    for (int c = 0; c < C; ++c) r2[c] = 0.0;
    for (int q = 0; q < d; ++q) {
      const double xv = a.XT[(size_t)q * np + j];
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const double df = xcs[c * d + q] - xv;
        r2[c] = fma(df, df, r2[c]);
      }
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
      double k = 0.0, dk = 0.0;
      if (j < n && c < a.mc) cov_k_dk<KIND>(r2[c], a.variance, k, dk);
      a.Kx[(size_t)c * np + j] = k;
      a.Dk[(size_t)c * np + j] = dk;
    }
  }
  sk_fence_proxy_async();      // Kx is read back through the TMA engine (async proxy) by other CTAs
  sk_stamp(a.dbg, 1);
  grid.sync();
  sk_stamp(a.dbg, 2);

  const long long u0 = (long long)g * T / G, u1 = (long long)(g + 1) * T / G;
  const int nu = (int)(u1 - u0);
  const uint64_t pol_stream = sk_policy_evict_first(), pol_keep = sk_policy_evict_last();
  int it = 0;                  // chunks consumed so far by this CTA, over both passes: stage = it % STAGES, parity = (it / STAGES) & 1

  // ---- P1: Z = M k*  (row panels; warp w owns rows 4 w .. 4 w + 3 of the panel, lane l four columns of the chunk) ----
  if (want_var && nu > 0) {
    // decode u0 -> (panel p, column block cb); a second cursor (ip, icb) runs SK_STAGES chunks ahead and feeds the ring
    int p = 0;
    {
      long long bi = (long long)((sqrt(1.0 + 2.0 * (double)u0) - 1.0) * 0.5);
      while (bi > 0 && 2 * bi * (bi + 1) > u0) --bi;
      while (2 * (bi + 1) * (bi + 2) <= u0) ++bi;
      const long long rem = u0 - 2 * bi * (bi + 1);
      p = (int)(4 * bi + rem / (bi + 1));
    }
    int cb = (int)(u0 - sk_prefix1(p));
    int ip = p, icb = cb, issued = 0;
    auto feed1 = [&](int slot_it) {   // warp 0: one chunk of M (32 rows of 1 KB) and the matching 1 KB of every right-hand side
      const int st = slot_it % SK_STAGES;
      if (tid == 0) {
        const uint32_t bar = sk_smem_u32(&full_bar[st]);
        double *dstM = ring + (size_t)st * STAGE_DOUBLES;
        sk_mbar_expect_tx(bar, (uint32_t)(STAGE_DOUBLES * sizeof(double)));
        sk_tma_load_2d(sk_smem_u32(dstM), &mapM, icb * TILE, ip * SK_PH, bar, pol_stream);            // 32 rows x 128 columns of M
        sk_tma_load_2d(sk_smem_u32(dstM + SK_MTILE), &mapK, icb * TILE, 0, bar, pol_keep);            // the same columns of the C right-hand sides
      }
      if (++icb > (ip >> 2)) {
        ++ip;
        icb = 0;
      }
      ++issued;
    };
    if (tid == 0) sk_fence_proxy_async();
    for (int i = 0; i < SK_STAGES && i < nu; ++i) feed1(it + i);
    double acc[4][C];
#pragma unroll
    for (int rr = 0; rr < 4; ++rr)
#pragma unroll
      for (int c = 0; c < C; ++c) acc[rr][c] = 0.0;
    auto flush1 = [&](int pp) {
      const int slot = g - sk_owner(sk_prefix1(pp), G, T);
#pragma unroll
      for (int rr = 0; rr < 4; ++rr)
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const double v = warp_sum(acc[rr][c]);
          if (lane == 0) a.part1[(((size_t)pp * SK_SLOTS1 + slot) * C + c) * SK_PH + 4 * warp + rr] = v;
          acc[rr][c] = 0.0;
        }
    };
    for (int i = 0; i < nu; ++i, ++it) {
      const int st = it % SK_STAGES;
      sk_mbar_wait(sk_smem_u32(&full_bar[st]), (uint32_t)((it / SK_STAGES) & 1));
      const double *tile = ring + (size_t)st * STAGE_DOUBLES;
      // a lane's four columns are {2 l, 2 l + 1, 64 + 2 l, 64 + 2 l + 1}: each 16-byte shared load of the warp then covers 512
      // contiguous bytes (with columns 4 l .. 4 l + 3 the lanes stride 32 bytes and every load is a two-way bank conflict, ncu r2o)
      const double *mrow = tile + (4 * warp) * TILE + 2 * lane;
      double2 m[4][2];
#pragma unroll
      for (int rr = 0; rr < 4; ++rr) {
        m[rr][0] = *reinterpret_cast<const double2 *>(mrow + rr * TILE);
        m[rr][1] = *reinterpret_cast<const double2 *>(mrow + rr * TILE + 64);
      }
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const double2 b0 = *reinterpret_cast<const double2 *>(tile + SK_MTILE + c * TILE + 2 * lane);
        const double2 b1 = *reinterpret_cast<const double2 *>(tile + SK_MTILE + c * TILE + 2 * lane + 64);
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
          double sacc = acc[rr][c];
          sacc = fma(m[rr][0].x, b0.x, sacc);
          sacc = fma(m[rr][0].y, b0.y, sacc);
          sacc = fma(m[rr][1].x, b1.x, sacc);
          sacc = fma(m[rr][1].y, b1.y, sacc);
          acc[rr][c] = sacc;
        }
      }
      if (++cb > (p >> 2)) {
        flush1(p);
        ++p;
        cb = 0;
      }
      __syncthreads();                                 // every warp has read stage st: refill it
      if (issued < nu) {
        if (tid == 0) sk_fence_proxy_async_smem();
        feed1(it + SK_STAGES);
      }
    }
    if (cb > 0) flush1(p);
  }
  sk_stamp(a.dbg, 3);
  if (want_var) {
    grid.sync();
    sk_stamp(a.dbg, 4);
    // Z_c = the partial sums of the (at most SK_SLOTS1) CTAs that shared a panel, added in slot order
    for (int e = g * SK_THREADS + tid; e < C * np; e += G * SK_THREADS) {
      const int c = e / np, i = e - c * np;
      a.Z[e] = sk_zsum<C>(a, c, i, G, T);
    }
    sk_stamp(a.dbg, 5);
    grid.sync();
    sk_stamp(a.dbg, 6);
  }

  // ---- P2: U = M^T Z  (column blocks; lane l owns four columns, warp w rows 4 w .. 4 w + 3 of every panel) --------
  if (want_dvar && nu > 0) {
    int cb = 0;
    {
      // prefix2(cb) = cb P - 2 cb (cb - 1) is increasing in cb for cb <= nb: search
      int lo = 0, hi = nb - 1;
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (sk_prefix2(mid, P) <= u0) lo = mid; else hi = mid - 1;
      }
      cb = lo;
    }
    int p = 4 * cb + (int)(u0 - sk_prefix2(cb, P));
    int ip = p, icb = cb, issued = 0;
    auto feed2 = [&](int slot_it) {
      const int st = slot_it % SK_STAGES;
      if (tid == 0) {
        const uint32_t bar = sk_smem_u32(&full_bar[st]);
        sk_mbar_expect_tx(bar, (uint32_t)(SK_MTILE * sizeof(double)));
        sk_tma_load_2d(sk_smem_u32(ring + (size_t)st * STAGE_DOUBLES), &mapM, icb * TILE, ip * SK_PH, bar, pol_stream);
      }
      if (++ip == P) {
        ++icb;
        ip = 4 * icb;
      }
      ++issued;
    };
    if (tid == 0) sk_fence_proxy_async();
    for (int i = 0; i < SK_STAGES && i < nu; ++i) feed2(it + i);
    double acc[C][4];
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[c][k] = 0.0;
    auto flush2 = [&](int cc) {
      // every warp parks its sums, then each thread adds the eight copies of its elements in warp order (a fixed order) -- two
      // barriers per flush; the first version took turns (nine barriers), which the CTAs at the short end of the triangle, with a
      // flush every few chunks, paid with 40 us of imbalance at 8 candidates
      const int slot = g - sk_owner(sk_prefix2(cc, P), G, T);
#pragma unroll
      for (int c = 0; c < C; ++c) {
        double *dst = wred + ((size_t)warp * C + c) * TILE + 2 * lane;
        *reinterpret_cast<double2 *>(dst) = make_double2(acc[c][0], acc[c][1]);
        *reinterpret_cast<double2 *>(dst + 64) = make_double2(acc[c][2], acc[c][3]);
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[c][k] = 0.0;
      }
      __syncthreads();
      for (int e = tid; e < C * TILE; e += SK_THREADS) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < SK_WARPS; ++w) v += wred[(size_t)w * C * TILE + e];
        const int c = e / TILE, r = e - c * TILE;
        a.part2[(((size_t)cc * SK_SLOTS2 + slot) * C + c) * TILE + r] = v;
      }
      __syncthreads();
    };
    // Z of the first panel; afterwards the next panel's values are fetched one chunk ahead
    if (tid < C * SK_PH) zs[0][tid] = a.Z[(size_t)(tid >> 5) * np + p * SK_PH + (tid & 31)];
    __syncthreads();
    for (int i = 0; i < nu; ++i, ++it) {
      const int st = it % SK_STAGES;
      int pn = p + 1, cbn = cb;                         // the chunk after this one
      if (pn == P) {
        ++cbn;
        pn = 4 * cbn;
      }
      double znext = 0.0;
      if (i + 1 < nu && tid < C * SK_PH) znext = a.Z[(size_t)(tid >> 5) * np + pn * SK_PH + (tid & 31)];
      sk_mbar_wait(sk_smem_u32(&full_bar[st]), (uint32_t)((it / SK_STAGES) & 1));
      const double *mrow = ring + (size_t)st * STAGE_DOUBLES + (4 * warp) * TILE + 2 * lane;
      const double *zc = zs[i & 1];
#pragma unroll
      for (int rr = 0; rr < 4; ++rr) {
        const double2 m0 = *reinterpret_cast<const double2 *>(mrow + rr * TILE);             // columns 2 l, 2 l + 1
        const double2 m1 = *reinterpret_cast<const double2 *>(mrow + rr * TILE + 64);        // columns 64 + 2 l, 64 + 2 l + 1
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const double z = zc[c * SK_PH + 4 * warp + rr];
          acc[c][0] = fma(m0.x, z, acc[c][0]);
          acc[c][1] = fma(m0.y, z, acc[c][1]);
          acc[c][2] = fma(m1.x, z, acc[c][2]);
          acc[c][3] = fma(m1.y, z, acc[c][3]);
        }
      }
      if (tid < C * SK_PH) zs[(i + 1) & 1][tid] = znext;
      const bool col_done = (p + 1 == P);
      __syncthreads();                                  // stage st and zs[i & 1] are free again
      if (issued < nu) {
        if (tid == 0) sk_fence_proxy_async_smem();
        feed2(it + SK_STAGES);
      }
      if (col_done) flush2(cb);
      p = pn;
      cb = cbn;
    }
    if (p > 4 * cb) flush2(cb);
  }
  sk_stamp(a.dbg, 7);
  if (want_dvar) {
    grid.sync();
    // U_c = the partial sums of the (at most SK_SLOTS2) CTAs that shared a column block, added in slot order
    for (int e = g * SK_THREADS + tid; e < C * np; e += G * SK_THREADS) {
      const int c = e / np, j = e - c * np;
      a.U[e] = sk_usum<C>(a, c, j, G, T, P);
    }
    grid.sync();
  }
  sk_stamp(a.dbg, 8);

  // ---- P3: the reductions over the training points; part3[g][c][2 + 2 d] -----------------------------------------------------
  // Every CTA takes n / G consecutive training points.  Work items (candidate c, block of SK_QC input dimensions) are dealt to the
  // warps, eight per round, so that all candidates are reduced at once; inside a warp the lanes stride over the CTA's points, park
  // their partial sums in shared memory and one thread per sum adds them in lane order.
  const int K3 = 2 + 2 * d;
  {
    const int nq = want_g ? (d + SK_QC - 1) / SK_QC : 1;
  for (int round0 = 0; round0 < a.mc * nq; round0 += SK_WARPS) {      // eight work items per round, one per warp
    const int item = round0 + warp;
    if (item < a.mc * nq) {
      const int c = item / nq, q0 = (item - c * nq) * SK_QC;
      double a_mu = 0.0, a_vv = 0.0, g1[SK_QC], g2[SK_QC];
#pragma unroll
      for (int k = 0; k < SK_QC; ++k) g1[k] = g2[k] = 0.0;
      {
        // every CTA takes an equal contiguous share of the training points; a lane takes four of them per round and issues ALL their
        // loads (index clamped, value masked afterwards) before the first multiply-add
        const int share = (n + G - 1) / G, jlo = g * share, jhi = min(n, jlo + share);
        for (int jb = jlo; jb < jhi; jb += 128) {
          double al[4], kx[4], zz[4], dk[4], uu[4], xt[4][SK_QC];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int j = jb + lane + 32 * u;
            const int jj = j < jhi ? j : jlo;
            al[u] = a.alpha[jj];
            kx[u] = a.Kx[(size_t)c * np + jj];
            zz[u] = a.Z[(size_t)c * np + jj];
            dk[u] = a.Dk[(size_t)c * np + jj];
            uu[u] = a.U[(size_t)c * np + jj];
#pragma unroll
            for (int k = 0; k < SK_QC; ++k) xt[u][k] = a.XT[(size_t)min(q0 + k, d - 1) * np + jj];
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const bool ok = jb + lane + 32 * u < jhi;
            const double alv = ok ? al[u] : 0.0;
            if (q0 == 0) {
              a_mu = fma(kx[u], alv, a_mu);
              const double z = (ok && want_var) ? zz[u] : 0.0;
              a_vv = fma(z, z, a_vv);
            }
            const double dkv = (ok && want_g) ? dk[u] : 0.0;
            const double w1 = dkv * alv;
            const double w2 = want_dvar ? dkv * uu[u] : 0.0;
#pragma unroll
            for (int k = 0; k < SK_QC; ++k) {
              const double df = (q0 + k < d) ? xcs[c * d + min(q0 + k, d - 1)] - xt[u][k] : 0.0;
              g1[k] = fma(w1, df, g1[k]);
              g2[k] = fma(w2, df, g2[k]);
            }
          }
        }
      }
      // lane partials -> shared memory (the chunk ring is idle now): [warp][value][lane], odd pitch
      double *mine = ring + (size_t)warp * (2 + 2 * SK_QC) * 33 + lane;
      mine[0] = a_mu;
      mine[33] = a_vv;
#pragma unroll
      for (int k = 0; k < SK_QC; ++k) {
        mine[(2 + k) * 33] = g1[k];
        mine[(2 + SK_QC + k) * 33] = g2[k];
      }
    }
    // one thread per (warp, value): the 32 lane partials added in lane order -- a fixed order, whatever the candidate count.
    // (A shuffle-tree version of this reduction took 23 us per work item for reasons the SASS does not show; this one takes 2.)
    __syncthreads();
    for (int e = tid; e < SK_WARPS * (2 + 2 * SK_QC); e += SK_THREADS) {
      const int w = e / (2 + 2 * SK_QC), i = e - w * (2 + 2 * SK_QC);
      const int item = round0 + w;
      if (item < a.mc * nq) {
        const int c = item / nq, q0 = (item - c * nq) * SK_QC;
        const double *src = ring + ((size_t)w * (2 + 2 * SK_QC) + i) * 33;
        double v = 0.0;
#pragma unroll
        for (int l = 0; l < 32; ++l) v += src[l];
        double *dst = a.part3 + ((size_t)g * C + c) * K3;
        if (i < 2) {
          if (q0 == 0) dst[i] = v;
        } else if (i < 2 + SK_QC) {
          if (q0 + i - 2 < d) dst[2 + q0 + i - 2] = v;
        } else {
          if (q0 + i - 2 - SK_QC < d) dst[2 + d + q0 + i - 2 - SK_QC] = v;
        }
      }
    }
    __syncthreads();
  }
  }
  sk_stamp(a.dbg, 9);
  // ---- the CTA that arrives last adds the per-CTA sums (CTA order, fixed) and applies the scalings ----------------------------
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(a.ticket, 1u) == (unsigned)(G - 1));
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int e = tid; e < a.mc * K3; e += SK_THREADS) {
    const int c = e / K3, i = e - c * K3;
    double v = 0.0;
    int gg = 0;
    for (; gg + 8 <= G; gg += 8) {                       // eight loads in flight, added in CTA order
      double t[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) t[k] = __ldcg(a.part3 + ((size_t)(gg + k) * C + c) * K3 + i);
#pragma unroll
      for (int k = 0; k < 8; ++k) v += t[k];
    }
    for (; gg < G; ++gg) v += __ldcg(a.part3 + ((size_t)gg * C + c) * K3 + i);
    if (i == 0) {
      a.mu[c] = v;
    } else if (i == 1) {
      if (want_var) a.var[c] = a.var_base - v;
    } else if (i < 2 + d) {
      if (want_g) a.dmu[(size_t)c * d + (i - 2)] = a.inv_ls[i - 2] * v;                     // gradients_X(alpha^T, X*, X) / l
    } else {
      if (want_dvar) a.dvar[(size_t)c * d + (i - 2 - d)] = -2.0 * a.inv_ls[i - 2 - d] * v;    // gradients_X(-2 Kx^T Ky^-1, X*, X) / l
    }
  }
  if (tid == 0) *a.ticket = 0u;                          // ready for the next call on this stream
  if (a.dbg && tid == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    a.dbg[10] = t;
  }
}

// scratch of the fused call beyond Kx / Dk (8 x np doubles each):
//   part12: part1 | part2 = np / 32 * SLOTS1 * 8 * 32 + np / 128 * SLOTS2 * 8 * 128 = 128 np doubles
//   part3:  (148 * 4) CTAs x 8 candidates x (2 + 2 d) doubles, followed by 8 doubles whose first word is the arrival counter
size_t skinny_fused_part12_doubles(int np) { return (size_t)np * (SK_SLOTS1 * 8 + SK_SLOTS2 * 8); }
size_t skinny_fused_part3_doubles(int d) { return (size_t)(148 * 4) * 8 * (2 + 2 * d) + 8; }

typedef CUresult (*SkEncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static SkEncodeTiledFn sk_encode_fn() {
  static const SkEncodeTiledFn fn = [] {
    void *ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      return (SkEncodeTiledFn)ptr;
    return (SkEncodeTiledFn) nullptr;
  }();
  return fn;
}
// rows x cols fp64 matrix with row pitch ld doubles -> 2-d tensor map with boxes of box_rows x 128 doubles (no swizzle: a warp reads a
// 1 KB row of the box with consecutive 32-byte pieces per lane, which is conflict free as it stands)
static int sk_make_map(CUtensorMap *map, const double *base, int rows, int cols, int ld, int box_rows) {
  SkEncodeTiledFn fn = sk_encode_fn();
  GPB_REQUIRE(fn != nullptr, "skinny: cuTensorMapEncodeTiled is not available");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(double)};
  cuuint32_t box[2] = {(cuuint32_t)TILE, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void *)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GPB_REQUIRE(r == CUDA_SUCCESS, "skinny: cuTensorMapEncodeTiled failed (%d) for %d x %d, ld %d", (int)r, rows, cols, ld);
  return 0;
}

template <int KIND, int C>
static int launch_skinny_fused_t(SkinnyParams &a, double *part12, double *part3, cudaStream_t s) {
  static std::atomic<int> per_sm[64];                    // co-resident CTAs per SM of this instantiation, per device (0 = not yet known)
  static std::atomic<int> sms[64];
  static FuncConfigMask configured{0};
  constexpr size_t smem = ((size_t)SK_STAGES * (SK_MTILE + C * TILE) + (size_t)SK_WARPS * C * TILE) * sizeof(double) + 128;
  int dev = 0;
  GPB_CUDA(cudaGetDevice(&dev));
  GPB_REQUIRE(dev >= 0 && dev < 64, "skinny: device ordinal %d out of range", dev);
  {
    FuncConfigOnce once(configured);
    if (once.needed) {
      int occ = 0, nsm = 0, coop = 0;
      GPB_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
      GPB_REQUIRE(coop != 0, "skinny: the device does not support cooperative launches");
      GPB_CUDA(cudaFuncSetAttribute(skinny_fused_kernel<KIND, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      GPB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, skinny_fused_kernel<KIND, C>, SK_THREADS, smem));
      GPB_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
      GPB_REQUIRE(occ >= 1 && nsm >= 1, "skinny: kernel does not fit an SM");
      sms[dev].store(nsm, std::memory_order_release);
      per_sm[dev].store(std::min(occ, 4), std::memory_order_release);
    }
  }
  CUtensorMap mapM, mapK;
  GPB_TRY(sk_make_map(&mapM, a.M, a.np, a.np, a.np, SK_PH));
  GPB_TRY(sk_make_map(&mapK, a.Kx, C, a.np, a.np, C));
  const int nb = a.np / TILE, P = a.np / SK_PH;
  const long long T = 2ll * nb * (nb + 1);
  // G <= 4 (nb + 1) keeps the chunks of one row panel within SK_SLOTS1 CTAs and those of one column block within SK_SLOTS2;
  // G <= T gives every CTA at least one chunk, so that the CTAs sharing a panel / column block are consecutive (no unwritten slot)
  int G = std::min(sms[dev].load() * per_sm[dev].load(), 4 * (nb + 1));
  G = (int)std::min<long long>(std::min(G, 148 * 4), T);
  a.part1 = part12;
  a.part2 = a.part1 + (size_t)P * SK_SLOTS1 * 8 * SK_PH;
  a.part3 = part3;
  a.ticket = reinterpret_cast<unsigned int *>(a.part3 + (size_t)(148 * 4) * 8 * (2 + 2 * a.d));
  static const bool dbg_on = [] { const char *e = getenv("GPB_SKINNY_DBG"); return e && atoi(e) != 0; }();
  static unsigned long long *dbg_dev = nullptr;
  a.dbg = nullptr;
  if (dbg_on) {
    if (!dbg_dev) GPB_CUDA(cudaMalloc(&dbg_dev, 16 * sizeof(unsigned long long)));
    GPB_CUDA(cudaMemsetAsync(dbg_dev, 0, 16 * sizeof(unsigned long long), s));
    a.dbg = dbg_dev;
  }
  void *args[] = {&mapM, &mapK, &a};
  GPB_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void *>(skinny_fused_kernel<KIND, C>), dim3(G), dim3(SK_THREADS), args, smem, s));
  count_launch();
  if (dbg_on) {       // phase times of CTA 0 (ns): P0, sync, P1, sync, Zfin, sync, P2, sync, P3, final
    unsigned long long h[16];
    GPB_CUDA(cudaStreamSynchronize(s));
    GPB_CUDA(cudaMemcpy(h, dbg_dev, sizeof(h), cudaMemcpyDeviceToHost));
    fprintf(stderr, "skinny C=%d G=%d np=%d:", C, G, a.np);
    for (int k = 1; k <= 10; ++k) fprintf(stderr, " %lld", h[k] >= h[k - 1] && h[k - 1] ? (long long)(h[k] - h[k - 1]) : -1ll);
    fprintf(stderr, "\n");
  }
  return 0;
}

// mc <= 8 candidates (raw coordinates Xc on the device).  Kx, Dk, Z, U: 8 x np doubles each; part12 / part3 as sized above, the last 8
// doubles of part3 (the arrival counter) zeroed once.  level as in predict_block (1, 2 or 3).
int launch_skinny_fused(int kind, const double *M, int np, int n, int d, int mc, int level, const double *XT, const double *Xc,
                        const double *ls, const double *inv_ls, const double *alpha, double variance, double var_base, double *Kx,
                        double *Dk, double *Z, double *U, double *part12, double *part3, double *mu, double *var, double *dmu,
                        double *dvar, cudaStream_t s) {
  GPB_REQUIRE(mc >= 1 && mc <= 8 && d >= 1 && d <= 64 && (level == 1 || level == 2 || level == 3), "skinny: bad arguments");
  SkinnyParams a;
  a.M = M; a.XT = XT; a.Xc = Xc; a.ls = ls; a.inv_ls = inv_ls; a.alpha = alpha;
  a.Kx = Kx; a.Dk = Dk; a.Z = Z; a.U = U;
  a.mu = mu; a.var = var; a.dmu = dmu; a.dvar = dvar;
  a.variance = variance; a.var_base = var_base;
  a.np = np; a.n = n; a.d = d; a.mc = mc; a.level = level;
  // one instantiation per candidate count: the row-panel pass costs in proportion to it once it leaves the HBM bound (5 = the
  // anchors of GPyOpt's AcquisitionOptimizer refined in lockstep)
#define GPB_SKF(K_)                                                            \
  do {                                                                         \
    if (mc == 1) return launch_skinny_fused_t<K_, 1>(a, part12, part3, s);     \
    if (mc == 2) return launch_skinny_fused_t<K_, 2>(a, part12, part3, s);     \
    if (mc == 3) return launch_skinny_fused_t<K_, 3>(a, part12, part3, s);     \
    if (mc == 4) return launch_skinny_fused_t<K_, 4>(a, part12, part3, s);     \
    if (mc == 5) return launch_skinny_fused_t<K_, 5>(a, part12, part3, s);     \
    if (mc == 6) return launch_skinny_fused_t<K_, 6>(a, part12, part3, s);     \
    return launch_skinny_fused_t<K_, 8>(a, part12, part3, s);                  \
  } while (0)
  if (kind == GPB_KERN_RBF) GPB_SKF(GPB_KERN_RBF);
  GPB_SKF(GPB_KERN_MATERN52);
#undef GPB_SKF
}

}  // namespace gpb
