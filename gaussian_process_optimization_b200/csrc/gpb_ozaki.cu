// fp64 products on the INT8 tensor cores (tcgen05.mma kind::i8, accumulators in TMEM, operands by TMA): an Ozaki-scheme GEMM
// engine for the large products of the cholinv recursion (gpb_chol.cu).  EXPERIMENTAL, off by default (GPB_OZAKI_MIN_N /
// gpb_set_ozaki): the default path of libgpb200 stays the fp64 DMMA engine of gpb_gemm.cu.
//
// Why: on sm_100a the fp64 tensor path is the warp-level DMMA (64 FMA/clk/SM = 37 TFLOP/s, shared with DFMA) -- tcgen05 has no
// f64 kind -- while the int8 kind runs at ~4.5 POP/s.  Every fp64 operand is cut, row by row, into S balanced radix-256 digits
//     a_ik = 2^ea_i * sum_s A_s[i][k] 2^(1 - 8 s),   A_s in [-128, 127]
// (error-free up to the rounding of the last digit: a_ik / 2^ea_i, |.| < 1/2, becomes the 64-bit integer q = rint(. 2^(8 S - 1)),
// whose bytes are taken from the low end with carries), and the product is the sum over digit pairs of EXACT integer products
//     C_ij = 2^(ea_i + eb_j) * sum_w 2^(2 - 8 w) * sum_{s + t = w} (A_s B_t^T)_ij,       w = 2 .. S + 1  (pairs with s + t > S + 1
// fall below the last digit and are dropped; balanced digits make them zero-mean), S (S + 1) / 2 int8 products in all.
// S = 7 (28 products, 53 bits relative to the largest entry of a row) keeps a whole NLL + gradient evaluation within 6e-11 /
// 4e-10 of LAPACK at cond(Ky) ~ 1e8 and 1e-13 in ordinary cases; S = 8 is indistinguishable from the fp64 recursion
// (scripts/ozaki_numerics_study.py -> profiles/r1i_ozaki_numerics_study.json).
//
// One kernel does the whole product for a 256 x 256 output tile: for each weight w (smallest first) the digit pairs s + t = w
// are chained along k into ONE int32 accumulation in TMEM (two M = 128 accumulators = all 512 columns; at most 1023 k-blocks per
// accumulation: 128^2 * 128 * 1023 < 2^31), drained by the epilogue warps: every drain but the last
// parks its exact int32 sums in a private coalesced scratch plane, the last one combines them in fp64 (smallest terms first) and
// applies the row / column scales, alpha and beta.  The 256 x 256 tile is what the operand
// traffic asks for: with 128 x 256 tiles and a double-buffered accumulator the kernel was bound by the L2 -> shared-memory fill
// (90 B/clk/SM; 16.0 ms for 8192^3 against 10.7 ms with the loads switched off; DESIGN.md section 3, "What bounds it").
// Triangular operands are exploited as in gpb_gemm.cu (per-tile k-ranges at 128 granularity, lower tiles only).
//
// CTA pairs (cluster of 2): two tiles with a common k-range fetch half of the shared operand tile each and multicast it into both
// shared memories; the MMA warps release a stage in both CTAs (tcgen05.commit ... multicast::cluster).
//
// Modular mode (S = 10 .. 18 planes = moduli instead of digits; gpb_crt.cuh): the planes hold the balanced residues of the scaled integer
// operand modulo pairwise coprime p_i <= 256, the weight loop runs over the moduli with ONE product each (16 instead of 28), every drain
// parks its sums reduced modulo p_i as int8, and oz_crt_combine_kernel rebuilds the integer product by the Chinese remainder theorem.
//
// Warp roles (320 threads): warp 0 = TMA producer (one lane), warp 1 = TMEM allocation + MMA issue (one lane),
// warps 2-9 = epilogue (TMEM lane quarter = warp % 4, accumulator = (warp - 2) / 4).  3-stage smem ring of {A digit tile
// 256 x 128 B, B digit tile 256 x 128 B}, 128-byte swizzle, K-major.
#include <cuda.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "gpb_common.cuh"
#include "gpb_crt.cuh"

namespace gpb {

namespace oz {
constexpr int BM = 256, BN = 256, BKB = 128;   // output tile of a CTA (two M = 128 accumulators); bytes (= int8 elements) of k per stage
constexpr int STAGES = 3;
constexpr int A_BYTES = BM * BKB, B_BYTES = BN * BKB, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /* alignment slack */ + 256 /* barriers */;
constexpr int PREFETCH = 6;                     // k-blocks of L2 prefetch distance in the TMA producer
constexpr int NTHREADS = 320;                  // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue
constexpr int GROUP_KB = 1023;                 // k-blocks per int32 accumulation: 128^2 * 128 * 1023 = 2.145e9 < 2^31
constexpr int MAX_S = 8;                        // 8 S - 1 <= 63: the scaled operand must fit a 64-bit integer
// Modular (CRT) mode: a plane count S in [crt::MINMOD, crt::MAXMOD] means "S moduli" instead of "S digits" (gpb_crt.cuh): ONE int8
// product per modulus, every drain parks the balanced residues of its int32 sums as int8, oz_crt_combine_kernel rebuilds the
// integer product (Garner + Horner) and writes C.  16 moduli carry 56 bits per operand at k = 16384 (7 digits: 55), 17 carry 59.
__constant__ crt::ModTable c_mods = crt::make_table();

struct Params {
  int M, N, K, S;
  int tri_out, klo_mode, khi_mode;
  double alpha, beta;
  const double *ra, *rb;   // 2^ea_i (M), 2^eb_j (N)
  double *C; int ldc;
  uint4 *P; size_t plane;  // int32 parking space of the drains, in uint4: [drain][tile][epilogue warp 8][chunk 8][j 8][lane 32] --
                           // every warp store / load is one contiguous 512-byte line (the owner lane reads back what it wrote)
  int tiles_m, tiles_n;
  int cl, share_a;         // CTA pairs (cluster of 2) that share one operand tile by TMA multicast: A (pair along n) or B (pair along m)
  int nmod;                // > 0: modular mode with that many moduli (S = nmod planes per operand)
  int dbg;                 // measurement only: 1 = no TMA loads (MMA-bound rate), 2 = no epilogue memory traffic, 3 = no drains but the last (wrong results)
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
// Bounded wait: a pipeline bug traps (the launch fails with an error) instead of hanging the device.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 8000000000ll) __trap();
  }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_mc(uint32_t dst, const CUtensorMap *map, int c0, int c1, int c2, uint32_t bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%2, %3, %4}], [%5], %6;" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
               : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap *map, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
      "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major operand tile, 128-byte swizzle: rows of 128 bytes, 8-row groups 1024 bytes apart (SBO = 64 x 16 B), descriptor
// version 1 (Blackwell), layout type 2 (SWIZZLE_128B).  The tile base is 1024-byte aligned; a k-step of 32 bytes inside the
// swizzle atom advances the start-address field by 2.
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// kind::i8: D = s32 (bits 4-5 = 2), A = B = signed 8 bit (bits 7-9, 10-12 = 1), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

__global__ void __launch_bounds__(NTHREADS, 1)
ozaki_mma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const __grid_constant__ CUtensorMap mapH,
                 const Params p) {
  // ---- tile coordinates ----
  // CTAs start in blockIdx order, one per SM, and a tile's duration is proportional to its k-range (2 .. K / 128 blocks with the
  // triangular modes): enumerate the tiles GLOBALLY longest first, so that the launch ends on its shortest tiles (the grouped
  // order of the first version left a 64-block tile to start last: triangular products took 0.75 of a full one instead of 0.52).
  // CTA pairs (p.cl == 2, a cluster of two consecutive CTAs): two tiles with the same k-range -- neighbours along n when the range
  // depends on the row (they share the A tile), along m when it depends on the column (they share the B tile).  Each CTA fetches
  // half of the shared tile and multicasts it into both shared memories: 48 KB instead of 64 KB of L2 reads per CTA and step.
  int tm, tn;
  const int rank = (p.cl == 2) ? (int)(blockIdx.x & 1) : 0;
  {
    const int t = (p.cl == 2) ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int gm = (p.cl == 2 && !p.share_a) ? p.tiles_m / 2 : p.tiles_m;   // grid of tiles (or of tile pairs)
    const int gn = (p.cl == 2 && p.share_a) ? p.tiles_n / 2 : p.tiles_n;
    if (p.khi_mode == 2) {                 // k <= column block: column by column from the right
      tn = gn - 1 - t / gm;
      tm = t % gm;
    } else if (p.khi_mode == 1) {          // k <= row block: row by row from the bottom
      tm = gm - 1 - t / gn;
      tn = t % gn;
    } else if (p.klo_mode == 2) {          // k >= column block: column by column from the left
      tn = t / gm;
      tm = t % gm;
    } else if (p.klo_mode == 1) {          // k >= row block: row by row from the top
      tm = t / gn;
      tn = t % gn;
    } else {                               // equal lengths: groups of 4 tile rows share their B panels in L2
      constexpr int GROUP = 4;
      const int in_group = GROUP * gn;
      const int gid = t / in_group, first = gid * GROUP;
      const int gsz = min(gm - first, GROUP);
      const int r = t - gid * in_group;
      tm = first + r % gsz;
      tn = r / gsz;
    }
    if (p.cl == 2) {
      if (p.share_a) tn = 2 * tn + rank;
      else tm = 2 * tm + rank;
    }
  }
  const int row0 = tm * BM, col0 = tn * BN;
  const int row_last = min(row0 + BM, p.M) - 128;          // first row / column of the last 128-block of the tile
  const int col_last = min(col0 + BN, p.N) - 128;
  if (p.tri_out && col0 > row_last + 127) return;          // tile entirely above the diagonal
  // pair mode: is the partner tile active?  (lower-tile outputs always pair along n: the partner is the other column tile)
  bool dual = p.cl == 2;
  if (dual && p.tri_out) dual = ((tn ^ 1) * BN <= row_last + 127);
  const uint16_t pair_mask = 3;
  int klo = (p.klo_mode == 1) ? row0 : (p.klo_mode == 2) ? col0 : 0;
  int khi = (p.khi_mode == 1) ? row_last + 128 : (p.khi_mode == 2) ? col_last + 128 : p.K;
  if (khi > p.K) khi = p.K;
  const int kb0 = klo / BKB, nkb = (khi - klo) / BKB;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = base + STAGES * STAGE_BYTES;       // full[STAGES], empty[STAGES], tfull, tempty, tmem slot
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
  const uint32_t tfull_bar = bars + 8u * (2 * STAGES), tempty_bar = bars + 8u * (2 * STAGES + 1);
  const uint32_t tmem_slot = bars + 8u * (2 * STAGES + 2);
  volatile uint32_t *tmem_slot_ptr = (volatile uint32_t *)(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), dual ? 2 : 1);               // pair mode: a stage is free when BOTH MMA warps are done with it
    }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, NTHREADS / 32 - 2);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (dual) cluster_sync_all();                            // the partner's barriers exist before anything is multicast to them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  // accumulations per tile: one per digit weight w = S + 1 .. 2 (the pairs s + t = w chained along k), or one per modulus
  const int nW = p.nmod ? p.nmod : p.S;
  auto weight_range = [&](int wi, int &w, int &s_lo, int &s_hi) {
    if (p.nmod) { w = 2 * (wi + 1); s_lo = s_hi = wi + 1; }
    else { w = p.S + 1 - wi; s_lo = max(1, w - p.S); s_hi = min(p.S, w - 1); }
  };
  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int wi = 0; wi < nW; ++wi) {
        int w, s_lo, s_hi;
        weight_range(wi, w, s_lo, s_hi);
        for (int s = s_lo; s <= s_hi; ++s) {
          const int t = w - s;
          for (int kb = 0; kb < nkb; ++kb) {
            mbar_wait(empty_bar(stage), phase ^ 1);
            if (p.dbg == 1) {
              mbar_arrive(full_bar(stage));
            } else {
              mbar_expect_tx(full_bar(stage), STAGE_BYTES);
              const uint32_t dst = base + stage * STAGE_BYTES;
              if (!dual) {
                tma_load_3d(dst, &mapA, (kb0 + kb) * BKB, row0, s - 1, full_bar(stage));
                tma_load_3d(dst + A_BYTES, &mapB, (kb0 + kb) * BKB, col0, t - 1, full_bar(stage));
              } else if (p.share_a) {      // own B tile; half of the common A tile to both CTAs
                tma_load_3d(dst + A_BYTES, &mapB, (kb0 + kb) * BKB, col0, t - 1, full_bar(stage));
                tma_load_3d_mc(dst + rank * (A_BYTES / 2), &mapH, (kb0 + kb) * BKB, row0 + rank * 128, s - 1, full_bar(stage), pair_mask);
              } else {                     // own A tile; half of the common B tile to both CTAs
                tma_load_3d(dst, &mapA, (kb0 + kb) * BKB, row0, s - 1, full_bar(stage));
                tma_load_3d_mc(dst + A_BYTES + rank * (B_BYTES / 2), &mapH, (kb0 + kb) * BKB, col0 + rank * 128, t - 1, full_bar(stage), pair_mask);
              }
              // pull the tiles PREFETCH k-blocks ahead into L2: the 3-stage ring alone cannot cover a DRAM miss
              if (kb + PREFETCH < nkb) {
                tma_prefetch_3d(&mapA, (kb0 + kb + PREFETCH) * BKB, row0, s - 1);
                tma_prefetch_3d(&mapB, (kb0 + kb + PREFETCH) * BKB, col0, t - 1);
              }
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issue: two M = 128 accumulators (rows 0-127 -> TMEM columns 0-255, rows 128-255 -> columns 256-511) =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int wi = 0; wi < nW; ++wi) {
        int w, s_lo, s_hi;
        weight_range(wi, w, s_lo, s_hi);
        const int npairs = s_hi - s_lo + 1;
        const int total = npairs * nkb;
        for (int g0 = 0; g0 < total; g0 += GROUP_KB) {
          const int g1 = min(total, g0 + GROUP_KB);
          const bool skip_drain = p.dbg == 3 && !((wi == nW - 1) && (g0 + GROUP_KB >= total));   // measurement only: no hand-over
          if (!(p.dbg == 3 && (wi != 0 || g0 != 0))) {
            mbar_wait(tempty_bar, acc_phase ^ 1);
            tc_fence_after();
          }
          for (int idx = g0; idx < g1; ++idx) {
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            const uint32_t a_addr = base + stage * STAGE_BYTES, b_addr = a_addr + A_BYTES;
            const uint64_t ad = smem_desc(a_addr), bd = smem_desc(b_addr);
#pragma unroll
            for (int k = 0; k < BKB / 32; ++k) {
              const uint32_t accum = (idx > g0 || k > 0) ? 1u : 0u;
              umma_i8(tmem_base, ad + 2 * k, bd + 2 * k, IDESC, accum);
              umma_i8(tmem_base + BN, ad + (128 * BKB / 16) + 2 * k, bd + 2 * k, IDESC, accum);
            }
            if (dual) umma_commit_mc(empty_bar(stage), pair_mask);
            else umma_commit(empty_bar(stage));
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          if (!skip_drain) {
            umma_commit(tfull_bar);
            acc_phase ^= 1;
          }
        }
      }
    }
  } else {
    // ===== epilogue (8 warps): TMEM -> fp64 running sum (global, L2 resident) -> final scaling =====
    const int q = warp & 3;                               // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;                     // which accumulator (row half of the tile)
    const int row = row0 + half * 128 + q * 32 + lane;
    const bool row_ok = row < p.M;
    uint32_t acc_phase = 0;
    int drain = 0;
    uint4 *Pw = p.P + ((size_t)(tm * p.tiles_n + tn) * 8 + (warp - 2)) * (8 * 8 * 32) + lane;
    double *Crow = p.C + (size_t)(row_ok ? row : 0) * p.ldc + col0;
    const double ra = row_ok ? p.alpha * p.ra[row] : 0.0;
    int ncols = min(BN, p.N - col0);
    if (p.tri_out) ncols = min(ncols, row0 + half * 128 + 128 - col0);   // 128-blocks above the diagonal are left untouched
    if (!row_ok) ncols = 0;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * BN);
    for (int wi = 0; wi < nW; ++wi) {
      int w, s_lo, s_hi;
      weight_range(wi, w, s_lo, s_hi);
      const int npairs = s_hi - s_lo + 1;
      const int total = npairs * nkb;
      for (int g0 = 0; g0 < total; g0 += GROUP_KB) {
        const bool last = !p.nmod && (wi == nW - 1) && (g0 + GROUP_KB >= total);
        if (p.dbg == 3 && !last) { ++drain; continue; }
        mbar_wait(tfull_bar, acc_phase);
        tc_fence_after();
        if (p.nmod) {
          // modular mode: one accumulation per modulus (the host guarantees nkb <= GROUP_KB); park the balanced residues of the
          // exact int32 sums as int8, 16 columns per uint4, [modulus][tile][epilogue warp 8][chunk 16][lane 32]
          const crt::ModParams mp = c_mods.m[wi];
          uint4 *dst = p.P + (size_t)wi * p.plane + ((size_t)(tm * p.tiles_n + tn) * 8 + (warp - 2)) * (16 * 32) + lane;
          for (int c0 = 0; c0 < ncols; c0 += 64) {
            uint32_t v[32], v2[32];
            tmem_ld32(taddr + c0, v);
            tmem_ld32(taddr + c0 + 32, v2);
            tmem_ld_wait();
            uint32_t wd[16];
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              wd[j / 4] = crt::pack4(crt::residue_of_sum((int)v[j], mp), crt::residue_of_sum((int)v[j + 1], mp),
                                     crt::residue_of_sum((int)v[j + 2], mp), crt::residue_of_sum((int)v[j + 3], mp));
              wd[8 + j / 4] = crt::pack4(crt::residue_of_sum((int)v2[j], mp), crt::residue_of_sum((int)v2[j + 1], mp),
                                         crt::residue_of_sum((int)v2[j + 2], mp), crt::residue_of_sum((int)v2[j + 3], mp));
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) dst[(c0 / 16 + u) * 32] = make_uint4(wd[4 * u], wd[4 * u + 1], wd[4 * u + 2], wd[4 * u + 3]);
          }
        } else if (!last) {
          // park the exact int32 sums of this drain (write only; the combination happens once, below)
          uint4 *dst = Pw + (size_t)drain * p.plane;
          for (int c0 = 0; c0 < ncols; c0 += 64) {             // ncols is a multiple of 128; two loads in flight per wait
            uint32_t v[32], v2[32];
            tmem_ld32(taddr + c0, v);
            tmem_ld32(taddr + c0 + 32, v2);
            tmem_ld_wait();
            if (p.dbg == 2) continue;
#pragma unroll
            for (int j = 0; j < 32; j += 4) dst[(c0 + j) * 8] = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);   // ((c0/32)*8 + j/4)*32
#pragma unroll
            for (int j = 0; j < 32; j += 4) dst[(c0 + 32 + j) * 8] = make_uint4(v2[j], v2[j + 1], v2[j + 2], v2[j + 3]);
          }
        } else {
          // C = alpha 2^(ea + eb) sum_drains 2^(2 - 8 w) P + beta C, smallest terms first.  The parked planes are read back with
          // all planes of 16 columns in flight at once (the combination is a chain of L2 round trips otherwise).
          constexpr int MAXP = 6;
          double scd[MAXP];
          int nd = 0;
          bool fast = true;
          for (int w2 = p.S + 1; w2 >= 2; --w2) {
            const int np2 = min(p.S, w2 - 1) - max(1, w2 - p.S) + 1;
            const int tot2 = np2 * nkb;
            for (int h0 = 0; h0 < tot2; h0 += GROUP_KB) {
              if (w2 == 2 && h0 + GROUP_KB >= tot2) break;         // the drain still in TMEM
              if (nd < MAXP) scd[nd] = exp2((double)(2 - 8 * w2));
              else fast = false;
              ++nd;
            }
          }
          const double sc_last = exp2((double)(2 - 8 * 2));
          for (int c0 = 0; c0 < ncols; c0 += 16) {
            uint32_t v[16];
            tmem_ld16(taddr + c0, v);                            // in flight under the plane reads below
            double acc[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j] = 0.0;
            if (fast) {
              uint4 u[MAXP][4];                                  // up to 6 planes x 16 columns in flight
#pragma unroll
              for (int d = 0; d < MAXP; ++d)
                if (d < nd) {
#pragma unroll
                  for (int j4 = 0; j4 < 4; ++j4) u[d][j4] = (Pw + (size_t)d * p.plane)[(c0 + 4 * j4) * 8];
                }
#pragma unroll
              for (int d = 0; d < MAXP; ++d)
                if (d < nd) {
#pragma unroll
                  for (int j4 = 0; j4 < 4; ++j4) {
                    const int j = 4 * j4;
                    acc[j] = fma(scd[d], (double)(int)u[d][j4].x, acc[j]);
                    acc[j + 1] = fma(scd[d], (double)(int)u[d][j4].y, acc[j + 1]);
                    acc[j + 2] = fma(scd[d], (double)(int)u[d][j4].z, acc[j + 2]);
                    acc[j + 3] = fma(scd[d], (double)(int)u[d][j4].w, acc[j + 3]);
                  }
                }
            } else {
              int d = 0;
              for (int w2 = p.S + 1; w2 >= 2; --w2) {
                const int np2 = min(p.S, w2 - 1) - max(1, w2 - p.S) + 1;
                const int tot2 = np2 * nkb;
                const double sc = exp2((double)(2 - 8 * w2));
                for (int h0 = 0; h0 < tot2; h0 += GROUP_KB) {
                  if (w2 == 2 && h0 + GROUP_KB >= tot2) break;
                  const uint4 *src = Pw + (size_t)d * p.plane;
#pragma unroll
                  for (int j = 0; j < 16; j += 4) {
                    const uint4 uu = src[(c0 + j) * 8];
                    acc[j] = fma(sc, (double)(int)uu.x, acc[j]);
                    acc[j + 1] = fma(sc, (double)(int)uu.y, acc[j + 1]);
                    acc[j + 2] = fma(sc, (double)(int)uu.z, acc[j + 2]);
                    acc[j + 3] = fma(sc, (double)(int)uu.w, acc[j + 3]);
                  }
                  ++d;
                }
              }
            }
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
              double2 o;
              o.x = fma(sc_last, (double)(int)v[j], acc[j]) * (ra * p.rb[col0 + c0 + j]);
              o.y = fma(sc_last, (double)(int)v[j + 1], acc[j + 1]) * (ra * p.rb[col0 + c0 + j + 1]);
              if (p.beta != 0.0) {
                const double2 c = *reinterpret_cast<double2 *>(Crow + c0 + j);
                o.x = fma(p.beta, c.x, o.x);
                o.y = fma(p.beta, c.y, o.y);
              }
              *reinterpret_cast<double2 *>(Crow + c0 + j) = o;
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar);
        acc_phase ^= 1;
        ++drain;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (dual) cluster_sync_all();                            // nobody leaves while the partner may still write into this CTA
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ---- digit extraction ------------------------------------------------------------------------------------------------------
// op(P)[i][k] = P[i * ld + k] (LAYOUT_ROWK) or P[k * ld + i] (LAYOUT_COLK); tri: 0 = full, 1 = valid where k-block <= i-block
// (lower-triangular factor, row-major), 2 = valid where k-block >= i-block (its transpose).  Invalid 128-blocks are never read
// (they hold scratch of the recursion) and get zero digits.
__device__ __forceinline__ bool block_valid(int tri, int ib, int kb) { return tri == 0 || (tri == 1 ? kb <= ib : kb >= ib); }

template <int LAYOUT>
__global__ void __launch_bounds__(256) oz_absmax_kernel(const double *__restrict__ P, int ld, int tri, unsigned long long *amax) {
  const int ib = blockIdx.x, kb = blockIdx.y;
  if (!block_valid(tri, ib, kb)) return;
  const int tid = threadIdx.x;
  if (LAYOUT == LAYOUT_ROWK) {
    const int tx = tid & 31, ty = tid >> 5;
    for (int rr = ty; rr < 128; rr += 8) {
      const double *row = P + (size_t)(ib * 128 + rr) * ld + kb * 128;
      double m = fmax(fmax(fabs(row[tx]), fabs(row[tx + 32])), fmax(fabs(row[tx + 64]), fabs(row[tx + 96])));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
      if (tx == 0) atomicMax(amax + ib * 128 + rr, (unsigned long long)__double_as_longlong(m));
    }
  } else {
    __shared__ double sm[128];
    const int i = tid & 127, half = tid >> 7;
    const double *col = P + (size_t)(kb * 128 + half * 64) * ld + ib * 128 + i;
    double m = 0.0;
#pragma unroll 8
    for (int k = 0; k < 64; ++k) m = fmax(m, fabs(col[(size_t)k * ld]));
    if (half == 1) sm[i] = m;
    __syncthreads();
    if (half == 0) atomicMax(amax + ib * 128 + i, (unsigned long long)__double_as_longlong(fmax(m, sm[i])));
  }
}

// One CTA per 32 (i) x 128 (k) sub-block: stage through shared memory, then every thread cuts 16 consecutive k of one row into S
// digits and stores them as one 16-byte vector per digit plane (the 8 threads of a row write one full 128-byte line per plane).
// digits: [S][R][K] int8, k contiguous.
template <int LAYOUT>
__global__ void __launch_bounds__(256)
oz_split_kernel(const double *__restrict__ P, int ld, int R, int K, int tri, int S, int nmod, int beta,
                const unsigned long long *__restrict__ amax, int8_t *__restrict__ digits, double *__restrict__ scale) {
  __shared__ double sm[32][129];
  const int i0 = blockIdx.x * 32, k0 = blockIdx.y * 128;
  const int tid = threadIdx.x;
  const bool valid = block_valid(tri, i0 >> 7, k0 >> 7);
  const int r = tid >> 3, seg = (tid & 7) * 16;
  const size_t plane = (size_t)R * K;
  int8_t *out = digits + (size_t)(i0 + r) * K + k0 + seg;
  if (!valid) {
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (int s = 0; s < S; ++s) *reinterpret_cast<uint4 *>(out + s * plane) = z;
    return;
  }
  if (LAYOUT == LAYOUT_ROWK) {
    const int c = tid & 127, rr0 = tid >> 7;
#pragma unroll 4
    for (int rr = rr0; rr < 32; rr += 2) sm[rr][c] = P[(size_t)(i0 + rr) * ld + k0 + c];
  } else {
    const int c = tid & 31, kk0 = tid >> 5;
#pragma unroll 4
    for (int kk = kk0; kk < 128; kk += 8) sm[c][kk] = P[(size_t)(k0 + kk) * ld + i0 + c];
  }
  __syncthreads();
  const double am = __longlong_as_double((long long)amax[i0 + r]);
  int e = 0;
  if (am > 0.0 && am < 1e308) frexp(am, &e);            // am = f 2^e, f in [0.5, 1)
  e += 1;                                               // |x| 2^-e < 1/2: the leading digit stays within [-64, 64] + carry
  if (nmod) {
    // modular mode: q = rint(x 2^(beta - e)), |q| <= 2^(beta - 1) <= 2^61; plane i holds the balanced residue q mod p_i
    if (seg == 0) scale[i0 + r] = ldexp(1.0, e - beta);
    const double upm = ldexp(1.0, beta - e);
    unsigned long long qm[16];                          // q + 2^62 > 0: no sign handling in the residues
#pragma unroll
    for (int j = 0; j < 16; ++j) qm[j] = crt::bias_operand(__double2ll_rn(sm[r][seg + j] * upm));
    for (int i = 0; i < nmod; ++i) {
      const crt::ModParams mp = c_mods.m[i];
      uint32_t wds[4];
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4)
        wds[q4] = crt::pack4(crt::residue_of_biased(qm[q4 * 4], mp), crt::residue_of_biased(qm[q4 * 4 + 1], mp),
                             crt::residue_of_biased(qm[q4 * 4 + 2], mp), crt::residue_of_biased(qm[q4 * 4 + 3], mp));
      *reinterpret_cast<uint4 *>(out + i * plane) = make_uint4(wds[0], wds[1], wds[2], wds[3]);
    }
    return;
  }
  if (seg == 0) scale[i0 + r] = ldexp(1.0, e);          // every valid sub-block of the row writes the same value
  const double up = ldexp(1.0, 8 * S - 1 - e);          // x -> q = rint(x 2^-e 2^(8 S - 1)), |q| < 2^(8 S - 2) <= 2^62
  long long q[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) q[j] = __double2ll_rn(sm[r][seg + j] * up);
  for (int s = S - 1; s >= 0; --s) {                    // low byte first; balanced: d in [-128, 127], carry into the next byte
    uint32_t wds[4];
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4) {
      uint32_t wv = 0;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        long long &v = q[q4 * 4 + b];
        const long long d = (s == 0) ? v : (((v + 128) & 255) - 128);
        v = (v - d) >> 8;
        wv |= ((uint32_t)(int)d & 0xFFu) << (8 * b);
      }
      wds[q4] = wv;
    }
    *reinterpret_cast<uint4 *>(out + s * plane) = make_uint4(wds[0], wds[1], wds[2], wds[3]);
  }
}

// ---- modular mode: reconstruction ----------------------------------------------------------------------------------------
// One thread per (row, 4 columns): reads one 32-bit word (four int8 residues) per modulus from the planes the drains parked (a warp
// reads 128 contiguous bytes per modulus), rebuilds the four integers (crt::reconstruct: Garner's mixed-radix digits from
// compile-time constants, Horner in fp64), applies the row / column scales, alpha and beta and writes 32 bytes of C.
template <int NMOD, bool PACKED>
__global__ void __launch_bounds__(256) oz_crt_combine_kernel(const Params p) {
  const unsigned g = blockIdx.x * 256u + threadIdx.x;
  const int j4 = g & 3, lane = (g >> 2) & 31, chunk = (g >> 7) & 15, w8 = (g >> 11) & 7, tile = (int)(g >> 14);
  const int tm = tile / p.tiles_n, tn = tile - tm * p.tiles_n;
  const int q = (w8 + 2) & 3, half = w8 >> 2;               // the epilogue warp w8 + 2 of ozaki_mma_kernel owns these rows
  const int row0 = tm * BM, col0 = tn * BN;
  const int row = row0 + half * 128 + q * 32 + lane;
  int ncols = min(BN, p.N - col0);
  if (p.tri_out) ncols = min(ncols, row0 + half * 128 + 128 - col0);
  if (row >= p.M || chunk * 16 >= ncols) return;
  const uint32_t *src = reinterpret_cast<const uint32_t *>(p.P + (((size_t)tile * 8 + w8) * 16 + chunk) * 32 + lane) + j4;
  uint32_t u[NMOD];
#pragma unroll
  for (int i = 0; i < NMOD; ++i) u[i] = src[(size_t)i * p.plane * 4];
  const int c = col0 + chunk * 16 + j4 * 4;
  const double ra = p.alpha * p.ra[row];
  double *Cp = p.C + (size_t)row * p.ldc + c;
  double o[4];
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    int r[NMOD];
#pragma unroll
    for (int i = 0; i < NMOD; ++i) r[i] = (int)(int8_t)((u[i] >> (8 * b)) & 0xFFu);
    o[b] = crt::reconstruct<NMOD, PACKED>(r) * (ra * p.rb[c + b]);
  }
  if (p.beta != 0.0) {
    const double2 c0 = *reinterpret_cast<const double2 *>(Cp), c1 = *reinterpret_cast<const double2 *>(Cp + 2);
    o[0] = fma(p.beta, c0.x, o[0]);
    o[1] = fma(p.beta, c0.y, o[1]);
    o[2] = fma(p.beta, c1.x, o[2]);
    o[3] = fma(p.beta, c1.y, o[3]);
  }
  *reinterpret_cast<double2 *>(Cp) = make_double2(o[0], o[1]);
  *reinterpret_cast<double2 *>(Cp + 2) = make_double2(o[2], o[3]);
}
// GPB_OZAKI_COMBINE=0: the Garner sums as one multiply-add per digit; 1 (default): as dp4a over packed digits (same integers)
template <int NMOD> static void launch_combine(const Params &p, cudaStream_t st) {
  static int packed = -1;
  if (packed < 0) { const char *e = getenv("GPB_OZAKI_COMBINE"); packed = e ? atoi(e) : 1; }
  if (packed) oz_crt_combine_kernel<NMOD, true><<<p.tiles_m * p.tiles_n * 64, 256, 0, st>>>(p);
  else oz_crt_combine_kernel<NMOD, false><<<p.tiles_m * p.tiles_n * 64, 256, 0, st>>>(p);
}

// ---- host side -----------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)ptr;
  }
  return fn;
}

// digits [S][R][K] int8 -> 3-d tensor map, box {128 bytes of k, box_rows, 1}, 128-byte swizzle
static int make_map(CUtensorMap *map, const int8_t *digits, int R, int K, int S, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  GPB_REQUIRE(fn != nullptr, "ozaki: cuTensorMapEncodeTiled is not available");
  cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)R, (cuuint64_t)S};
  cuuint64_t strides[2] = {(cuuint64_t)K, (cuuint64_t)R * (cuuint64_t)K};
  cuuint32_t box[3] = {(cuuint32_t)BKB, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void *)digits, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GPB_REQUIRE(r == CUDA_SUCCESS, "ozaki: cuTensorMapEncodeTiled failed (%d) for R=%d K=%d S=%d", (int)r, R, K, S);
  return 0;
}

struct Workspace {
  int8_t *dA = nullptr, *dB = nullptr;
  size_t capA = 0, capB = 0;
  double *ra = nullptr, *rb = nullptr;
  unsigned long long *amax = nullptr;   // [0, cap) for A, [cap, 2 cap) for B
  size_t cap_rows = 0;
  void *T = nullptr;                    // int32 planes of the drains
  size_t capT = 0;
  // digit planes of a B operand kept across calls (the predictive products multiply every candidate block with the same L^-1)
  struct Cached {
    int8_t *dig = nullptr; size_t cap = 0;
    double *scale = nullptr; size_t cap_scale = 0;
    const double *ptr = nullptr; int ld = 0, layout = 0, tri = 0, R = 0, K = 0, S = 0;
    unsigned long long epoch = 0;
  } cache[2];
};
static std::atomic<unsigned long long> g_epoch{1};
// One workspace per (device, stream): models driven from several host threads (concurrent restarts) run on streams of their own.
struct Slot { int dev; cudaStream_t st; Workspace *ws; };
static std::mutex g_ws_mutex;
static std::vector<Slot> g_slots;
static Workspace *workspace_for(int dev, cudaStream_t st) {
  std::lock_guard<std::mutex> lock(g_ws_mutex);
  for (auto &s : g_slots)
    if (s.dev == dev && s.st == st) return s.ws;
  g_slots.push_back(Slot{dev, st, new Workspace()});
  return g_slots.back().ws;
}

// The stream is being destroyed (its model is): give the planes back.  Without this every model that ever used the engine left
// several GB behind, keyed by a stream handle that no longer exists (20 models at N = 16384 exhausted the 180 GB, r2f).
static void release_workspace(Workspace *w) {
  for (void *q : {(void *)w->dA, (void *)w->dB, (void *)w->ra, (void *)w->rb, (void *)w->amax, w->T, (void *)w->cache[0].dig,
                  (void *)w->cache[0].scale, (void *)w->cache[1].dig, (void *)w->cache[1].scale})
    if (q) cudaFree(q);
  delete w;
}

static int ensure(void **p, size_t *cap, size_t need) {
  if (*cap >= need) return 0;
  if (*p) GPB_CUDA(cudaFree(*p));
  *p = nullptr;
  *cap = 0;
  GPB_CUDA(cudaMalloc(p, need));
  *cap = need;
  return 0;
}

static int g_slices = 8;   // 8 digits: indistinguishable from the fp64 engine in every parity test; 7 is 25% faster
static int g_min_n = -1;   // -1: read GPB_OZAKI_MIN_N once; 0: off
static bool valid_planes(int s) { return (s >= 1 && s <= MAX_S) || (s >= crt::MINMOD && s <= crt::MAXMOD); }

}  // namespace oz

void ozaki_release_stream(cudaStream_t st) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return;
  std::lock_guard<std::mutex> lock(oz::g_ws_mutex);
  for (size_t i = 0; i < oz::g_slots.size(); ++i) {
    if (oz::g_slots[i].dev == dev && oz::g_slots[i].st == st) {
      oz::release_workspace(oz::g_slots[i].ws);
      oz::g_slots.erase(oz::g_slots.begin() + i);
      return;
    }
  }
}

static thread_local int t_suppress = 0;   // > 0: this host thread's calls stay on the fp64 DMMA engine (fallback after a failed residual check)
void ozaki_suppress(int on) { t_suppress += on ? 1 : -1; }
int ozaki_min_n() {
  if (t_suppress > 0) return 0;
  if (oz::g_min_n < 0) {
    const char *e = getenv("GPB_OZAKI_MIN_N");
    oz::g_min_n = e ? atoi(e) : 0;
    const char *s = getenv("GPB_OZAKI_SLICES");
    if (s && oz::valid_planes(atoi(s))) oz::g_slices = atoi(s);
  }
  return oz::g_min_n;
}
int ozaki_predict_planes() { ozaki_min_n(); return oz::g_slices >= crt::MINMOD ? crt::MAXMOD : oz::MAX_S; }
int ozaki_configure(int min_n, int slices) {
  GPB_REQUIRE(min_n >= 0 && oz::valid_planes(slices), "ozaki: min_n >= 0 and 1 <= slices <= %d (digits) or %d <= slices <= %d (moduli)",
              oz::MAX_S, crt::MINMOD, crt::MAXMOD);
  oz::g_min_n = min_n;
  oz::g_slices = slices;
  return 0;
}

// C = alpha op(A) op(B)^T + beta C through the int8 tensor cores.  Same argument meaning as gemm_launch (gpb_gemm.cu); tri_a /
// tri_b: validity pattern of the stored operand (see oz_split_kernel).  M, N, K multiples of 128.
void ozaki_invalidate() { oz::g_epoch.fetch_add(1, std::memory_order_relaxed); }

// cache_b = 1 or 2: keep the digit planes of B in that slot and reuse them while (pointer, shape, layout) match and no
// factorisation has run since (ozaki_invalidate()).
int ozaki_gemm_launch(int layout_a, int layout_b, const GemmArgs &g, int tri_a, int tri_b, int slices, cudaStream_t st, int cache_b) {
  using namespace oz;
  const int S = slices > 0 ? slices : g_slices;
  GPB_REQUIRE(g.M % 128 == 0 && g.N % 128 == 0 && g.K % 128 == 0 && g.M > 0 && g.N > 0 && g.K > 0, "ozaki: sizes must be multiples of 128");
  GPB_REQUIRE(!g.tri_out || g.M == g.N, "ozaki: tri_out needs a square output");
  GPB_REQUIRE(valid_planes(S), "ozaki: bad digit / modulus count %d", S);
  const int nmod = S >= crt::MINMOD ? S : 0;
  const int beta_bits = nmod ? crt::operand_bits(nmod, g.K) : 0;
  GPB_REQUIRE(!nmod || g.K / BKB <= GROUP_KB, "ozaki: modular mode needs k <= %d", GROUP_KB * BKB);
  int dev = 0;
  GPB_CUDA(cudaGetDevice(&dev));
  Workspace &ws = *workspace_for(dev, st);
  const bool same = (g.A == g.B && g.lda == g.ldb && layout_a == layout_b && tri_a == tri_b && g.M == g.N);
  GPB_TRY(ensure((void **)&ws.dA, &ws.capA, (size_t)S * g.M * g.K));
  GPB_REQUIRE(cache_b >= 0 && cache_b <= 2 && !(cache_b && same), "ozaki: bad operand cache slot");
  if (!same && !cache_b) GPB_TRY(ensure((void **)&ws.dB, &ws.capB, (size_t)S * g.N * g.K));
  const size_t rows = (size_t)(g.M > g.N ? g.M : g.N);
  if (ws.cap_rows < rows) {
    if (ws.ra) { GPB_CUDA(cudaFree(ws.ra)); GPB_CUDA(cudaFree(ws.rb)); GPB_CUDA(cudaFree(ws.amax)); }
    ws.cap_rows = 0;
    GPB_CUDA(cudaMalloc((void **)&ws.ra, rows * sizeof(double)));
    GPB_CUDA(cudaMalloc((void **)&ws.rb, rows * sizeof(double)));
    GPB_CUDA(cudaMalloc((void **)&ws.amax, 2 * rows * sizeof(unsigned long long)));
    ws.cap_rows = rows;
  }
  // int32 planes for every drain but the last: drains of weight w = ceil(pairs(w) * (K / 128) / GROUP_KB) at most
  int drains = 0;
  for (int w = S + 1; w >= 2; --w) {
    const int npairs = (S < w - 1 ? S : w - 1) - (1 > w - S ? 1 : w - S) + 1;
    drains += (npairs * (g.K / BKB) + GROUP_KB - 1) / GROUP_KB;
  }
  // (modular mode: one int8 plane per modulus, a quarter of that per plane)
  const size_t plane = (size_t)((g.M + BM - 1) / BM) * ((g.N + BN - 1) / BN) * (nmod ? BM * BN / 16 : BM * BN / 4);   // uint4 per drain
  GPB_TRY(ensure((void **)&ws.T, &ws.capT, (size_t)(nmod ? nmod : (drains > 1 ? drains - 1 : 1)) * plane * sizeof(uint4)));
  GPB_CUDA(cudaMemsetAsync(ws.amax, 0, 2 * ws.cap_rows * sizeof(unsigned long long), st));
  auto split = [&](int layout, const double *P, int ld, int R, int tri, unsigned long long *amax, int8_t *dig, double *scale) -> int {
    dim3 g1(R / 128, g.K / 128), g2(R / 32, g.K / 128);
    if (layout == LAYOUT_ROWK) {
      oz_absmax_kernel<LAYOUT_ROWK><<<g1, 256, 0, st>>>(P, ld, tri, amax);
      oz_split_kernel<LAYOUT_ROWK><<<g2, 256, 0, st>>>(P, ld, R, g.K, tri, S, nmod, beta_bits, amax, dig, scale);
    } else {
      oz_absmax_kernel<LAYOUT_COLK><<<g1, 256, 0, st>>>(P, ld, tri, amax);
      oz_split_kernel<LAYOUT_COLK><<<g2, 256, 0, st>>>(P, ld, R, g.K, tri, S, nmod, beta_bits, amax, dig, scale);
    }
    count_launch(2);
    GPB_CHECK_LAUNCH();
    return 0;
  };
  GPB_TRY(split(layout_a, g.A, g.lda, g.M, tri_a, ws.amax, ws.dA, ws.ra));
  const int8_t *digB = same ? ws.dA : ws.dB;
  const double *scaleB = same ? ws.ra : ws.rb;
  if (cache_b) {
    Workspace::Cached &c = ws.cache[cache_b - 1];
    const unsigned long long epoch = g_epoch.load(std::memory_order_relaxed);
    const bool hit = c.ptr == g.B && c.ld == g.ldb && c.layout == layout_b && c.tri == tri_b && c.R == g.N && c.K == g.K && c.S == S &&
                     c.epoch == epoch;
    if (!hit) {
      c.ptr = nullptr;
      GPB_TRY(ensure((void **)&c.dig, &c.cap, (size_t)S * g.N * g.K));
      GPB_TRY(ensure((void **)&c.scale, &c.cap_scale, (size_t)g.N * sizeof(double)));
      GPB_TRY(split(layout_b, g.B, g.ldb, g.N, tri_b, ws.amax + ws.cap_rows, c.dig, c.scale));
      c.ptr = g.B; c.ld = g.ldb; c.layout = layout_b; c.tri = tri_b; c.R = g.N; c.K = g.K; c.S = S; c.epoch = epoch;
    }
    digB = c.dig;
    scaleB = c.scale;
  } else if (!same) {
    GPB_TRY(split(layout_b, g.B, g.ldb, g.N, tri_b, ws.amax + ws.cap_rows, ws.dB, ws.rb));
  }
  CUtensorMap mapA, mapB, mapH;
  GPB_TRY(make_map(&mapA, ws.dA, g.M, g.K, S, BM));
  GPB_TRY(make_map(&mapB, digB, g.N, g.K, S, BN));
  // CTA pairs: the k-range must be common to the pair -> along n (shared A) unless it depends on the column (shared B)
  static int use_pairs = -1;
  if (use_pairs < 0) { const char *e = getenv("GPB_OZAKI_PAIRS"); use_pairs = e ? atoi(e) : 1; }
  const int tiles_m = (g.M + BM - 1) / BM, tiles_n = (g.N + BN - 1) / BN;
  const int share_a = (g.khi_mode == 2 || g.klo_mode == 2) ? 0 : 1;
  const int cl = (use_pairs && (share_a ? tiles_n % 2 == 0 : tiles_m % 2 == 0) && !(g.tri_out && !share_a)) ? 2 : 1;
  if (share_a) GPB_TRY(make_map(&mapH, ws.dA, g.M, g.K, S, 128));
  else GPB_TRY(make_map(&mapH, digB, g.N, g.K, S, 128));
  Params p;
  p.M = g.M; p.N = g.N; p.K = g.K; p.S = S;
  p.tri_out = g.tri_out; p.klo_mode = g.klo_mode; p.khi_mode = g.khi_mode;
  p.alpha = g.alpha; p.beta = g.beta;
  p.ra = ws.ra; p.rb = scaleB;
  p.C = g.C; p.ldc = g.ldc; p.P = (uint4 *)ws.T; p.plane = plane;
  p.tiles_m = tiles_m; p.tiles_n = tiles_n;
  p.cl = cl; p.share_a = share_a;
  static int dbg = -1;
  if (dbg < 0) { const char *e = getenv("GPB_OZAKI_DBG"); dbg = e ? atoi(e) : 0; }
  p.dbg = (nmod && dbg == 3) ? 0 : dbg;
  p.nmod = nmod;
  static FuncConfigMask configured{0};
  FuncConfigOnce once_configured(configured);
  if (once_configured.needed)
    GPB_CUDA(cudaFuncSetAttribute(ozaki_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(p.tiles_m * p.tiles_n);
    cfg.blockDim = dim3(NTHREADS);
    cfg.dynamicSmemBytes = SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cl;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    GPB_CUDA(cudaLaunchKernelEx(&cfg, ozaki_mma_kernel, mapA, mapB, mapH, p));
  }
  count_launch();
  GPB_CHECK_LAUNCH();
  if (nmod) {
    switch (nmod) {
      case 10: launch_combine<10>(p, st); break;
      case 11: launch_combine<11>(p, st); break;
      case 12: launch_combine<12>(p, st); break;
      case 13: launch_combine<13>(p, st); break;
      case 14: launch_combine<14>(p, st); break;
      case 15: launch_combine<15>(p, st); break;
      case 16: launch_combine<16>(p, st); break;
      case 17: launch_combine<17>(p, st); break;
      default: launch_combine<18>(p, st); break;
    }
    count_launch();
    GPB_CHECK_LAUNCH();
  }
  return 0;
}

// ---- host restatement of the modular arithmetic (the same gpb_crt.cuh code compiled for the host): test hooks behind the C ABI,
// used by the CPU tests to hold it bit-identical to oracle/ozaki_emulation.py.  Nothing on the product path calls them.
int ozaki_crt_bits(int nmod, long long k) { return crt::operand_bits(nmod, k); }
int ozaki_crt_host_residues(const double *A, int rows, int k, int nmod, int beta, signed char *planes, double *scale) {
  GPB_REQUIRE(A && planes && scale && rows > 0 && k > 0 && nmod >= crt::MINMOD && nmod <= crt::MAXMOD && beta >= 1 && beta <= 62,
              "ozaki_crt_host_residues: bad argument");
  static const crt::ModTable tab = crt::make_table();
  for (int i = 0; i < rows; ++i) {
    double am = 0.0;
    for (int j = 0; j < k; ++j) am = fmax(am, fabs(A[(size_t)i * k + j]));
    int e = 0;
    if (am > 0.0 && am < 1e308) frexp(am, &e);
    e += 1;
    scale[i] = ldexp(1.0, e - beta);
    const double up = ldexp(1.0, beta - e);
    for (int j = 0; j < k; ++j) {
      const long long q = llrint(A[(size_t)i * k + j] * up);
      for (int m = 0; m < nmod; ++m) planes[((size_t)m * rows + i) * k + j] = (signed char)crt::residue_of(q, tab.m[m]);
    }
  }
  return 0;
}
template <int NMOD> static void host_combine(const int *sums, size_t count, double *X) {
  static const crt::ModTable tab = crt::make_table();
  for (size_t e = 0; e < count; ++e) {
    int r[NMOD];
    for (int m = 0; m < NMOD; ++m) r[m] = crt::residue_of_sum(sums[(size_t)m * count + e], tab.m[m]);
    X[e] = crt::reconstruct<NMOD, true>(r);
    // both forms of the Garner sums give the same integers; NaN marks a disagreement for the tests
    const double x2 = crt::reconstruct<NMOD, false>(r);
    if (x2 != X[e]) X[e] = nan("");
  }
}
// sums: [nmod][count] exact int32 accumulations of the residue products -> X[count] = the integer products, rounded to fp64
int ozaki_crt_host_combine(const int *sums, long long count, int nmod, double *X) {
  GPB_REQUIRE(sums && X && count > 0 && nmod >= crt::MINMOD && nmod <= crt::MAXMOD, "ozaki_crt_host_combine: bad argument");
  switch (nmod) {
    case 10: host_combine<10>(sums, (size_t)count, X); break;
    case 11: host_combine<11>(sums, (size_t)count, X); break;
    case 12: host_combine<12>(sums, (size_t)count, X); break;
    case 13: host_combine<13>(sums, (size_t)count, X); break;
    case 14: host_combine<14>(sums, (size_t)count, X); break;
    case 15: host_combine<15>(sums, (size_t)count, X); break;
    case 16: host_combine<16>(sums, (size_t)count, X); break;
    case 17: host_combine<17>(sums, (size_t)count, X); break;
    default: host_combine<18>(sums, (size_t)count, X); break;
  }
  return 0;
}

}  // namespace gpb
