// fp64 tensor-core GEMM engine for the blocked Cholesky / inverse / predictive solves (north_star (b), (c)).
//
//   C = alpha * op(A) * op(B) + beta * C        (row-major everywhere, M,N multiples of 128, K multiple of 16)
//
// Blackwell has no f64 kind for tcgen05/UMMA, so the FP64 tensor path on sm_100a is the warp-level
// mma.sync.aligned.m8n8k4.f64 (SASS: DMMA.8x8x4, 16 issue cycles per sub-partition -> 64 FMA/clk/SM).  Operand tiles
// (rows x 16 k) are staged through shared memory with a 3-stage cp.async (LDGSTS) pipeline; row strides are padded
// (20 / rows+4 doubles == 4 mod 16) so that every half-warp's 64-bit fragment loads hit 16 distinct bank pairs (conflict
// free) in both storage orders.
//
// Three tile configurations share one kernel template (picked per launch from the number of output tiles):
//   BIG    64 x 128 per CTA, 4 warps (each 64 x 32), 2 CTAs resident per SM: while one CTA sits at its per-k-tile barrier
//          the other keeps the DMMA pipe busy (a single 128x128 CTA per SM left ~10% barrier/ramp bubbles in ncu);
//   MID    64 x 64, 4 warps (32 x 32)   -- 4x the CTAs of a 128-tile grid for problems that would not fill 148 SMs;
//   SMALL  32 x 32, 4 warps (16 x 16)   -- 16x the CTAs, for the bottom levels of the recursion (128..512-sized blocks).
//
// Triangular structure is exploited at 128-block granularity through a per-tile k-range (klo_mode / khi_mode) and through
// tri_out (only tiles touching the lower triangle are produced): operands that are triangular carry explicit zeros inside
// their 128x128 diagonal blocks, blocks strictly above the diagonal are never read.
#include <stdlib.h>

#include <vector>

#include "gpb_common.cuh"
#include "gpb_gemm_tile.cuh"

namespace gpb {

bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char *e = getenv("GPB_PDL");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on == 1;
}

template <int LA, int LB, int BM, int BN, int WARPS_M, int WARPS_N, int MIN_BLOCKS, int BK, int STAGES>
__global__ void __launch_bounds__(WARPS_M *WARPS_N * 32, MIN_BLOCKS) gemm_dmma_kernel(GemmArgs p) {
  constexpr int THREADS = WARPS_M * WARPS_N * 32;
  constexpr int WTM = BM / WARPS_M, WTN = BN / WARPS_N;  // warp tile
  constexpr int MI = WTM / 8, NI = WTN / 8;
  constexpr int A_TILE = tile_doubles<LA, BM, BK>();
  constexpr int B_TILE = tile_doubles<LB, BN, BK>();
  extern __shared__ __align__(16) double smem[];
  double *sA = smem;
  double *sB = smem + STAGES * A_TILE;

  pdl_trigger();
  // ---- tile coordinates ----
  int tm, tn;
  const int t = blockIdx.x;
  if (p.tri_out) {
    // Lower 128-blocks are enumerated band by band: a band is G consecutive block rows, walked as G x G squares from the left
    // (then the triangle on the diagonal), so that the ~148 blocks resident at a time share 2 G operand panels instead of a
    // whole block row's worth (row-major order re-read every column panel once per block row: 74 GB of DRAM reads for
    // Ky^-1 = M^T M at N = 16384 against 3 GB of operands).  Bands run top to bottom, which keeps "longest k-range first" for
    // klo_mode 1 (k >= row block).
    constexpr int R = 128 / BM, CW = 128 / BN;
    constexpr int per = R * CW;   // tiles of one 128x128 block
    // 6 x 6 squares: the resident CTAs then share 12 operand panels per square and DRAM reads of Ky^-1 = M^T M at N = 16384 are
    // 16.2 GB (3.2 GB of operands); 12 x 12 squares, one wave of CTAs per square, read 34.0 GB, whole block rows 74 GB
    // (profiles/r2m_traffic_band.json).  Same duration either way: the kernel is tensor-pipe bound.
    const int G = p.band > 0 ? p.band : 6;
    const int nb = p.M / 128;
    const int q = t / per, w_in = t - q * per;
    int i = (int)((sqrt(8.0 * (double)q + 1.0) - 1.0) * 0.5);   // block row of q in row-major triangular order
    while ((i + 1) * (i + 2) / 2 <= q) ++i;
    while (i * (i + 1) / 2 > q) --i;
    const int b0 = (i / G) * G;                                 // first block row of the band
    const int gb = min(G, nb - b0);                             // block rows in the band
    int qq = q - b0 * (b0 + 1) / 2;                             // index inside the band
    int bi, bj;
    if (qq < gb * b0) {                                         // rectangular part left of the diagonal squares
      const int cg = qq / (gb * G), w = qq - cg * (gb * G);
      bi = b0 + w % gb;
      bj = cg * G + w / gb;
    } else {                                                    // triangle on the diagonal
      qq -= gb * b0;
      int r = (int)((sqrt(8.0 * (double)qq + 1.0) - 1.0) * 0.5);
      while ((r + 1) * (r + 2) / 2 <= qq) ++r;
      while (r * (r + 1) / 2 > qq) --r;
      bi = b0 + r;
      bj = b0 + qq - r * (r + 1) / 2;
    }
    tm = bi * R + w_in / CW;
    tn = bj * CW + w_in % CW;
  } else {
    // grouped rasterisation: GROUP consecutive tile rows share their B panels while they are L2 resident
    const int tiles_m = p.M / BM, tiles_n = p.N / BN;
    constexpr int GROUP = 16;
    const int in_group = GROUP * tiles_n;
    const int gid = t / in_group;
    const int first = gid * GROUP;
    const int gsz = min(tiles_m - first, GROUP);
    const int r = t - gid * in_group;
    tm = first + r % gsz;
    tn = r / gsz;
    // longest k-ranges first: with a triangular k-range the tiles differ in length by up to K / 16 : 1, and a launch that
    // ends on its longest tiles leaves most SMs idle for a full-length tile (~1 ms at K = 8192)
    if (p.khi_mode == 2) tn = tiles_n - 1 - tn;   // k <= column block: right-most columns are the longest
    if (p.khi_mode == 1) tm = tiles_m - 1 - tm;   // k <= row block: bottom rows are the longest
  }
  const int row0 = tm * BM, col0 = tn * BN;
  const int rblk = row0 & ~127, cblk = col0 & ~127;  // enclosing 128-block
  int klo = (p.klo_mode == 1) ? rblk : (p.klo_mode == 2) ? cblk : 0;
  int khi = (p.khi_mode == 1) ? rblk + 128 : (p.khi_mode == 2) ? cblk + 128 : p.K;
  if (khi > p.K) khi = p.K;
  const int ktiles = (khi - klo) / BK;

  const int tid = threadIdx.x;
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

  pdl_wait();   // everything above is arithmetic on the launch arguments; the operands belong to the previous kernels
  // ---- prologue ----
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
    __syncthreads();
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

  // ---- epilogue: lane holds (row g, cols 2 tq, 2 tq + 1) of every 8x8 accumulator ----
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
}

// ---- optional per-launch timing (bench.py's roofline leg): CUDA events on the launching stream around every GEMM ----
struct GemmProfile {
  bool on = false;
  std::vector<cudaEvent_t> ev;  // pairs (start, stop)
  size_t used = 0;
  double flops = 0.0;           // executed tile flops (diagonal blocks counted in full)
  long long launches = 0;
};
static GemmProfile g_prof;
static int g_forced_config = 0;  // 0 auto, 1 BIG, 2 MID, 3 SMALL (tuning / tests)

int gemm_profile_enable(int on) {
  g_prof.on = on != 0;
  g_prof.used = 0;
  g_prof.flops = 0.0;
  g_prof.launches = 0;
  return 0;
}

static int g_big_config = 9;     // configuration of launches with >= 20 output blocks of 128x128 (100 + cfg selects it)

bool gemm_profile_is_on() { return g_prof.on; }

int gemm_force_config(int cfg) {
  if (cfg >= 100)
    g_big_config = cfg - 100;
  else
    g_forced_config = cfg;
  return 0;
}

// -> total milliseconds spent in GEMM launches since enable / last collect, executed flops, launch count
int gemm_profile_collect(double *ms, double *flops, long long *launches) {
  double total = 0.0;
  for (size_t i = 0; i + 1 < g_prof.used; i += 2) {
    GPB_CUDA(cudaEventSynchronize(g_prof.ev[i + 1]));
    float t = 0.f;
    GPB_CUDA(cudaEventElapsedTime(&t, g_prof.ev[i], g_prof.ev[i + 1]));
    total += t;
  }
  if (ms) *ms = total;
  if (flops) *flops = g_prof.flops;
  if (launches) *launches = g_prof.launches;
  g_prof.used = 0;
  g_prof.flops = 0.0;
  g_prof.launches = 0;
  return 0;
}

// -> milliseconds and executed flops of the LAST GEMM launch recorded since enable / collect (in an NLL+grad evaluation that
// is Ky^-1 = M^T M, the largest single launch); call before gemm_profile_collect.
static double g_last_flops = 0.0;
int gemm_profile_last(double *ms, double *flops) {
  if (g_prof.used < 2) {
    if (ms) *ms = 0.0;
    if (flops) *flops = 0.0;
    return 0;
  }
  GPB_CUDA(cudaEventSynchronize(g_prof.ev[g_prof.used - 1]));
  float t = 0.f;
  GPB_CUDA(cudaEventElapsedTime(&t, g_prof.ev[g_prof.used - 2], g_prof.ev[g_prof.used - 1]));
  if (ms) *ms = t;
  if (flops) *flops = g_last_flops;
  return 0;
}

static int prof_event(cudaStream_t s) {
  if (g_prof.used == g_prof.ev.size()) {
    cudaEvent_t e;
    GPB_CUDA(cudaEventCreate(&e));
    g_prof.ev.push_back(e);
  }
  GPB_CUDA(cudaEventRecord(g_prof.ev[g_prof.used++], s));
  return 0;
}

static double block_flops(const GemmArgs &g) {
  // sum over computed 128x128 blocks of 2 * 128 * 128 * (khi - klo)
  const int tm = g.M / 128, tn = g.N / 128;
  double k_sum = 0.0;
  for (int i = 0; i < tm; ++i) {
    const int jmax = g.tri_out ? i + 1 : tn;
    for (int j = 0; j < jmax; ++j) {
      const int row0 = i * 128, col0 = j * 128;
      int klo = (g.klo_mode == 1) ? row0 : (g.klo_mode == 2) ? col0 : 0;
      int khi = (g.khi_mode == 1) ? row0 + 128 : (g.khi_mode == 2) ? col0 + 128 : g.K;
      if (khi > g.K) khi = g.K;
      k_sum += (khi - klo);
    }
  }
  return 2.0 * 128 * 128 * k_sum;
}

template <int LA, int LB, int BM, int BN, int WARPS_M, int WARPS_N, int MIN_BLOCKS, int BK, int STAGES>
static int launch_t(const GemmArgs &g, cudaStream_t s) {
  constexpr int A_TILE = tile_doubles<LA, BM, BK>();
  constexpr int B_TILE = tile_doubles<LB, BN, BK>();
  constexpr int THREADS = WARPS_M * WARPS_N * 32;
  const size_t smem = (size_t)STAGES * (A_TILE + B_TILE) * sizeof(double);
  static FuncConfigMask configured{0};
  FuncConfigOnce once_configured(configured);
  if (once_configured.needed) {
    GPB_CUDA(cudaFuncSetAttribute(gemm_dmma_kernel<LA, LB, BM, BN, WARPS_M, WARPS_N, MIN_BLOCKS, BK, STAGES>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  const int b128m = g.M / 128, b128n = g.N / 128;
  const int blocks128 = g.tri_out ? b128m * (b128m + 1) / 2 : b128m * b128n;
  const int tiles = blocks128 * (128 / BM) * (128 / BN);
  if (tiles == 0) return 0;
  if (g_prof.on) GPB_TRY(prof_event(s));
  GPB_CUDA(launch_pdl(gemm_dmma_kernel<LA, LB, BM, BN, WARPS_M, WARPS_N, MIN_BLOCKS, BK, STAGES>, dim3(tiles), dim3(THREADS), smem, s, g));
  count_launch();
  if (g_prof.on) {
    GPB_TRY(prof_event(s));
    g_last_flops = block_flops(g);
    g_prof.flops += g_last_flops;
    g_prof.launches += 1;
  }
  return 0;
}

template <int LA, int LB>
static int launch_cfg(const GemmArgs &g, cudaStream_t s) {
  const int b128m = g.M / 128, b128n = g.N / 128;
  const int blocks128 = g.tri_out ? b128m * (b128m + 1) / 2 : b128m * b128n;
  int cfg = g_forced_config;
  // 64x64 tiles with four CTAs (16 warps) per SM and a two-stage pipeline beat the 64x128 / two-CTA configuration in every
  // storage order (scripts/gemm_sweep.py: 35.4 / 34.9 / 35.0 vs 34.6 / 34.1 / 34.4 TFLOP/s at 8192^3; cuBLAS 36.2): with the
  // DMMA issue slot held 16 cycles per instruction, more resident warps hide the fragment-load latency better than bigger
  // register tiles do
  // Up to a few hundred output blocks the 32x32 tiles win although they load twice the operand fragments per DMMA: 1024^3
  // takes 83 us with them against 138 us with 64x64 tiles, which leave 148 SMs with 256 CTAs (scripts/gemm_mid_sweep.py).  Alone,
  // 64x64 is ahead again from 1536^3 on, but inside the factorisation (triangular k-ranges, a side-stream product running
  // underneath) the 256-block launches of the 2048 level are still faster with 32x32: NLL+grad at N = 16384 takes 136.9 ms with
  // a threshold of 20 blocks, 136.0 with 100, 135.6 with 260..400, 136.0 with 600 (GPB_BIG_MIN_BLOCKS).
  static int big_min_blocks = -1;
  if (big_min_blocks < 0) {
    const char *e = getenv("GPB_BIG_MIN_BLOCKS");
    big_min_blocks = e ? atoi(e) : 260;
  }
  if (cfg == 0) cfg = (blocks128 >= big_min_blocks) ? g_big_config : 3;
  if (cfg == 1 && g_forced_config == 0 && blocks128 < 148) cfg = 2;   // the 64x128 policy of the first version, kept selectable
  // both operands k-contiguous: BK = 32 with two stages halves the per-k-tile barriers (+1.8% at 8192^3); the strided
  // layouts are faster with BK = 16 and three stages (scripts/gemm_sweep.py)
  if (cfg == 1 && LA == LAYOUT_ROWK && LB == LAYOUT_ROWK && g.K % 32 == 0 && g_forced_config == 0)
    return launch_t<LA, LB, 64, 128, 1, 4, 2, 32, 2>(g, s);
  if (cfg == 1) return launch_t<LA, LB, 64, 128, 1, 4, 2, 16, 3>(g, s);
  if (cfg == 2) return launch_t<LA, LB, 64, 64, 2, 2, 3, 16, 3>(g, s);
  if (cfg == 4 && g.K % 32 == 0) return launch_t<LA, LB, 64, 128, 1, 4, 2, 32, 2>(g, s);   // 64x128, BK = 32, two stages
  if (cfg == 9) return launch_t<LA, LB, 64, 64, 2, 2, 4, 16, 2>(g, s);                      // 64x64, 4 CTAs / SM, two stages
  return launch_t<LA, LB, 32, 32, 2, 2, 4, 16, 3>(g, s);
}

int gemm_launch(int la, int lb, const GemmArgs &g_in, cudaStream_t s) {
  static const int band_env = [] { const char *e = getenv("GPB_TRI_BAND"); return e ? atoi(e) : 0; }();
  GemmArgs g = g_in;
  if (g.band == 0) g.band = band_env;
  GPB_REQUIRE(g.M % 128 == 0 && g.N % 128 == 0 && g.K % BK_MIN == 0, "gemm: M,N must be multiples of 128 and K of 16 (got %d %d %d)",
              g.M, g.N, g.K);
  GPB_REQUIRE((g.lda % 2) == 0 && (g.ldb % 2) == 0 && (g.ldc % 2) == 0, "gemm: leading dimensions must be even");
  GPB_REQUIRE(!g.tri_out || g.M == g.N, "gemm: tri_out needs a square output");
  GPB_REQUIRE((g.klo_mode == 0 && g.khi_mode == 0) || g.K % 128 == 0, "gemm: triangular k-ranges need K to be a multiple of 128");
  GPB_REQUIRE((((uintptr_t)g.A | (uintptr_t)g.B | (uintptr_t)g.C) & 15) == 0, "gemm: operands must be 16-byte aligned");
  if (la == LAYOUT_ROWK && lb == LAYOUT_ROWK) return launch_cfg<LAYOUT_ROWK, LAYOUT_ROWK>(g, s);
  if (la == LAYOUT_ROWK && lb == LAYOUT_COLK) return launch_cfg<LAYOUT_ROWK, LAYOUT_COLK>(g, s);
  if (la == LAYOUT_COLK && lb == LAYOUT_COLK) return launch_cfg<LAYOUT_COLK, LAYOUT_COLK>(g, s);
  if (la == LAYOUT_COLK && lb == LAYOUT_ROWK) return launch_cfg<LAYOUT_COLK, LAYOUT_ROWK>(g, s);
  set_error("gemm: bad layouts %d %d", la, lb);
  return -2;
}

}  // namespace gpb
