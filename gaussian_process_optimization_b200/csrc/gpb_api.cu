// extern "C" surface of libgpb200.so (declared in include/gpb200.h) and the resident-model bookkeeping behind it.
#include <stdarg.h>

#include <algorithm>
#include <cmath>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "gpb_common.cuh"
#include "gpb_kernels.cuh"

namespace gpb {

static thread_local char g_err[1024] = "";
std::atomic<long long> g_launches{0};

std::mutex &func_config_mutex() {
  static std::mutex mu;
  return mu;
}

// environment knob read once (thread-safe: C++11 static initialisation)
// NVTX range around every C-ABI call (header-only NVTX3: a no-op costing a pointer check unless a tool such as ncu --nvtx is attached)
struct ApiRange {
  explicit ApiRange(const char *name) { nvtxRangePushA(name); }
  ~ApiRange() { nvtxRangePop(); }
};
#define GPB_RANGE(name) ApiRange api_range_(name)

static int env_int(const char *name, int fallback) {
  const char *e = getenv(name);
  return e ? atoi(e) : fallback;
}

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---- tiny layout kernels ----------------------------------------------------------------------------------------------
// Y (n x p row-major) -> Yc[p][np] zero padded
__global__ void pack_cols_kernel(const double *__restrict__ Y, int n, int p, int np, double *__restrict__ Yc) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= p * np) return;
  const int pp = e / np, i = e - pp * np;
  Yc[e] = (i < n) ? Y[(size_t)i * p + pp] : 0.0;
}
// Ac[p][np] -> out (n x p row-major)
__global__ void unpack_cols_kernel(const double *__restrict__ Ac, int n, int p, int np, double *__restrict__ out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n * p) return;
  const int i = e / p, pp = e - i * p;
  out[e] = Ac[(size_t)pp * np + i];
}
// dst[i][j] = (j <= i) ? src[i][j] : 0      (i, j < n)
__global__ void tril_copy_kernel(const double *__restrict__ src, int lds, double *__restrict__ dst, int ldd, int n) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (j >= n) return;
  dst[(size_t)i * ldd + j] = (j <= i) ? src[(size_t)i * lds + j] : 0.0;
}
// dst[i][j] = src[max(i,j)][min(i,j)]   (full symmetric from the lower triangle; reads of the upper part are transposed)
__global__ void sym_copy_kernel(const double *__restrict__ src, int lds, double *__restrict__ dst, int ldd, int n) {
  __shared__ double t[32][33];
  const int bi = blockIdx.y, bj = blockIdx.x;
  const int tx = threadIdx.x, ty = threadIdx.y;
  if (bj <= bi) {
    for (int r = ty; r < 32; r += 8) {
      const int i = bi * 32 + r, j = bj * 32 + tx;
      if (i < n && j < n) {
        const double v = (j <= i) ? src[(size_t)i * lds + j] : src[(size_t)j * lds + i];
        dst[(size_t)i * ldd + j] = v;
      }
    }
  } else {
    // upper tile (bi, bj): read the lower tile (bj, bi) and transpose through shared memory
    for (int r = ty; r < 32; r += 8) {
      const int i = bj * 32 + r, j = bi * 32 + tx;
      t[r][tx] = (i < n && j < n) ? src[(size_t)i * lds + j] : 0.0;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
      const int i = bi * 32 + r, j = bj * 32 + tx;
      if (i < n && j < n) dst[(size_t)i * ldd + j] = t[tx][r];
    }
  }
}
// dL_dK[i][j] = 0.5 (sum_p a_ip a_jp - P Wi[max][min])      exact_gaussian_inference.py:70
__global__ void dl_dk_kernel(const double *__restrict__ W, int ldw, const double *__restrict__ alpha, int np, int p, int n,
                             double *__restrict__ dst, int ldd) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (j >= n) return;
  double aa = 0.0;
  for (int pp = 0; pp < p; ++pp) aa = fma(alpha[(size_t)pp * np + i], alpha[(size_t)pp * np + j], aa);
  const double w = (j <= i) ? W[(size_t)i * ldw + j] : W[(size_t)j * ldw + i];
  dst[(size_t)i * ldd + j] = 0.5 * (aa - (double)p * w);
}
// out = sum_e a[e] b[e], single block, fixed order
__global__ void dot_kernel(const double *__restrict__ a, const double *__restrict__ b, int n, double *out) {
  __shared__ double scratch[32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += 1024) acc = fma(a[i], b[i], acc);
  acc = block_sum<1024>(acc, scratch);
  if (threadIdx.x == 0) *out = acc;
}
// dst (np x np) = [src (n x n, lower triangle mirrored) 0; 0 I]
__global__ void pad_sym_kernel(const double *__restrict__ src, int lds, int n, double *__restrict__ dst, int np) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (j >= np) return;
  double v;
  if (i < n && j < n)
    v = (j <= i) ? src[(size_t)i * lds + j] : src[(size_t)j * lds + i];
  else
    v = (i == j) ? 1.0 : 0.0;
  dst[(size_t)i * np + j] = v;
}

// ---- RAII device temp ------------------------------------------------------------------------------------------------
// Temporary device memory of one API call, stream-ordered (cudaMallocAsync / cudaFreeAsync on the call's stream, pool kept
// warm): a cudaMalloc / cudaFree pair costs milliseconds and synchronises the device, which the stateless Kern / linalg entry
// points would otherwise pay several times per call.  AllocStream names the stream for the DevBufs of the current scope.
static thread_local cudaStream_t g_alloc_stream = 0;
struct AllocStream {
  cudaStream_t prev;
  explicit AllocStream(cudaStream_t s) : prev(g_alloc_stream) { g_alloc_stream = s; }
  ~AllocStream() { g_alloc_stream = prev; }
};
struct DevBuf {
  void *p = nullptr;
  cudaStream_t s = 0;
  bool pooled = false;   // small blocks come from the stream-ordered pool; large ones (N x N matrices) from cudaMalloc, which is
                         // faster than growing the pool by gigabytes and does not keep them reserved afterwards
  ~DevBuf() {
    if (!p) return;
    if (pooled)
      cudaFreeAsync(p, s);
    else
      cudaFree(p);
  }
  int alloc(size_t bytes) {
    if (bytes == 0) bytes = 8;
    static FuncConfigMask configured{0};
    FuncConfigOnce once_configured(configured);
    if (once_configured.needed) {          // once per device: keep freed blocks in the pool instead of returning them
      int dev = 0;
      cudaMemPool_t pool;
      if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
      }
    }
    s = g_alloc_stream;
    pooled = bytes <= (size_t)64 << 20;
    if (pooled)
      GPB_CUDA(cudaMallocAsync(&p, bytes, s));
    else
      GPB_CUDA(cudaMalloc(&p, bytes));
    return 0;
  }
  double *d() { return reinterpret_cast<double *>(p); }
};

static int to_device(DevBuf &buf, const double *src, size_t count, int dev, const double **out, cudaStream_t s) {
  if (dev) {
    *out = src;
    return 0;
  }
  GPB_TRY(buf.alloc(count * sizeof(double)));
  GPB_CUDA(cudaMemcpyAsync(buf.p, src, count * sizeof(double), cudaMemcpyHostToDevice, s));
  *out = buf.d();
  return 0;
}

// lengthscale (nls = d or 1) -> device arrays ls[d], inv_ls[d]
static int upload_ls(const double *ls, int nls, int d, double *ls_dev, double *inv_ls_dev, cudaStream_t s) {
  std::vector<double> h(2 * d);
  for (int q = 0; q < d; ++q) {
    h[q] = ls[nls == 1 ? 0 : q];
    h[d + q] = 1.0 / h[q];
  }
  GPB_CUDA(cudaMemcpyAsync(ls_dev, h.data(), d * sizeof(double), cudaMemcpyHostToDevice, s));
  GPB_CUDA(cudaMemcpyAsync(inv_ls_dev, h.data() + d, d * sizeof(double), cudaMemcpyHostToDevice, s));
  GPB_CUDA(cudaStreamSynchronize(s));  // h goes out of scope
  return 0;
}

// the same without uploads or a synchronisation: the lengthscales travel as a kernel argument (d <= 64)
struct LsArg {
  double v[64];
};
__global__ void set_ls_kernel(LsArg a, int d, double *__restrict__ ls_dev, double *__restrict__ inv_ls_dev) {
  const int q = threadIdx.x;
  if (q < d) {
    ls_dev[q] = a.v[q];
    inv_ls_dev[q] = 1.0 / a.v[q];
  }
}
static int set_ls(const double *ls, int nls, int d, double *ls_dev, double *inv_ls_dev, cudaStream_t s) {
  if (d > 64) return upload_ls(ls, nls, d, ls_dev, inv_ls_dev, s);
  LsArg a;
  for (int q = 0; q < 64; ++q) a.v[q] = q < d ? ls[nls == 1 ? 0 : q] : 1.0;
  set_ls_kernel<<<1, 64, 0, s>>>(a, d, ls_dev, inv_ls_dev);
  count_launch();
  GPB_CHECK_LAUNCH();
  return 0;
}

}  // namespace gpb

using namespace gpb;

// =====================================================================================================================
// model
// =====================================================================================================================
struct gpb_model {
  int device = 0;  // the CUDA device the model was created on; every call must be made with that device current
  int kind = 0, ard = 1, d = 0, p = 1, n_cap = 0, np_cap = 0, cb = 0, nls = 0;
  int n = 0, np = 0;
  double variance = 1.0, noise = 1.0, jitter = 0.0;
  std::vector<double> ls;
  bool have_data = false, scaled_valid = false, fitted = false, have_wi = false;
  bool l_valid = true;   // false after gpb_model_adopt_state without the L region: A does not hold the factor of this posterior
  // CUDA-graph replay of the NLL+grad launch sequence (small / medium N: an evaluation is a chain of tens to hundreds of short
  // kernels whose host-side launch cost is a good part of the wall time).  One executable graph per want_grad, valid for one
  // (n, np, configuration epoch); the hyper-parameters reach the kernels through theta_dev and the pinned block.
  struct FitGraph {
    cudaGraphExec_t exec = nullptr;
    int n = -1, np = -1, epoch = -1, seen = 0;
    long long launches = 0;
    bool l_pending = false, have_wi = false;
    int l_from = 0;
  } graphs[2];
  bool graph_failed = false, own_stream = false;
  cudaEvent_t entry_ev = nullptr;   // own_stream: orders the model's private non-blocking stream after the legacy default stream at call entry
  cudaEvent_t exit_ev = nullptr;    // ... and the legacy stream after ours at the exit of the entry points that do not synchronise
  double *theta_dev = nullptr;
  int wi_from = 0;  // > 0 (after gpb_model_append): the leading wi_from block of W holds the old Ky^-1, downdated; rows beyond are stale
  cudaStream_t stream = 0;
  void *ws = nullptr;
  bool own_ws = false;
  Factor f;
  // device arrays (carved from ws)
  double *X = nullptr, *XsT = nullptr, *Yc = nullptr, *alpha = nullptr, *z = nullptr, *ls_dev = nullptr, *inv_ls_dev = nullptr;
  double *scal = nullptr;  // [0] logdet [1] alpha.Y [2..] kgrad outputs (d + 2)
  double *gpart = nullptr;
  // candidate-block buffers
  double *Xc = nullptr, *XcT = nullptr, *KxT = nullptr, *Vt = nullptr, *Ut = nullptr, *mu = nullptr, *var = nullptr, *dmu = nullptr,
         *dvar = nullptr, *fbuf = nullptr, *dfbuf = nullptr, *sdbuf = nullptr, *dsbuf = nullptr;
  double *topv = nullptr;
  long long *topi = nullptr;
  double *sk_part3 = nullptr;   // per-CTA sums + arrival counter of the fused M <= 8 call (gpb_skinny.cu)
  // Second lane of candidate-block buffers + stream (allocated on the first multi-block device-resident scoring pass): block b + 1 runs
  // its covariance rows, residue extraction / small kernels underneath block b's GEMMs instead of behind them
  struct Lane {
    double *Xc, *XcT, *XcgT, *KxT, *Vt, *Ut, *mu, *var, *dmu, *dvar, *fbuf, *dfbuf, *sdbuf, *dsbuf;
    cudaStream_t stream;
  };
  void *lane_mem = nullptr;
  Lane lane1 = {};
  cudaStream_t lane_stream = nullptr;
  cudaEvent_t lane_ev[3] = {nullptr, nullptr, nullptr};   // [0] entry hand-over, [1] top-k chain, [2] lane-1 completion
  // int8 engine (experimental): every fit that used it is followed by a residual check of the solve; a failed check refits on the
  // fp64 DMMA engine, and the predictive products of that posterior stay there as well
  bool engine_used = false, engine_ok = true;
  double last_residual = -1.0;
  double *pinned = nullptr;  // host: [0, 256) results / hyper-parameters of a fit; [PIN_X, +512) staged candidates and
                             // [PIN_OUT, +PIN_OUT_DOUBLES) staged results of the M <= 8 calls (one synchronisation per call)
  struct Staged { double *dst; size_t off, count; };
  std::vector<Staged> staged;  // device -> pinned copies in flight: unpacked into the caller's arrays after the synchronisation
  size_t staged_used = 0;
  FactorOverlap *ov = nullptr;  // streams / events of the two-stream factorisation schedule
  // local-penalisation state (AcquisitionLP.update_batches): batch points and hammer-function parameters on the device
  // Gower mixed-variable kernel patch (stationary.py:116-135): per-dimension flags / inverse ranges and the coordinate copies
  // scaled for it (K uses them; every gradient keeps the Euclidean XsT / XcT, like the reference)
  bool gower = false;
  std::vector<double> g_flag, g_inv;   // host: 1.0 = discrete; 1 / range (continuous) or 1 (discrete)
  double *XgT = nullptr, *XcgT = nullptr, *gflag_dev = nullptr, *ginv_dev = nullptr;
  double *lp_buf = nullptr;     // [Xb (nb x d) | r (nb) | s (nb)], own allocation of lp_cap rows
  int lp_cap = 0, lp_nb = 0, lp_transform = 0;
};

// A model's buffers, streams and events belong to one device: refuse calls made while another device is current.
static int check_device(const gpb_model *m, const char *what) {
  int dev = -1;
  GPB_CUDA(cudaGetDevice(&dev));
  GPB_REQUIRE(dev == m->device, "%s: the model lives on CUDA device %d but device %d is current", what, m->device, dev);
  if (m->own_stream && m->entry_ev) {
    // The caller asked for the legacy default stream; the model runs on a NON-blocking stream of its own (a blocking one must not be
    // captured while any thread touches the legacy stream).  Work the caller queued on the legacy stream before this call (e.g. a
    // torch op producing an input tensor) is ordered in front of ours explicitly; every entry point synchronises its stream before
    // it returns, which orders the other direction.
    // (an idle legacy stream has nothing to order against: one query instead of an event pair on the hot path of the small-N BO loop,
    // which makes tens of thousands of these calls)
    const cudaError_t q = cudaStreamQuery(cudaStreamLegacy);
    if (q == cudaErrorNotReady) {
      GPB_CUDA(cudaEventRecord(m->entry_ev, cudaStreamLegacy));
      GPB_CUDA(cudaStreamWaitEvent(m->stream, m->entry_ev, 0));
    } else if (q != cudaSuccess) {
      GPB_CUDA(q);
    }
  }
  return 0;
}

// Exit of an entry point that returns WITHOUT synchronising (gpb_model_acq_topk_dev): the caller named the legacy default stream,
// so whatever it queues there next (a collective, a copy, an event) must come after our work on the private stream.
static int leave_stream_ordered(gpb_model *m) {
  if (m->own_stream && m->exit_ev) {
    GPB_CUDA(cudaEventRecord(m->exit_ev, m->stream));
    GPB_CUDA(cudaStreamWaitEvent(cudaStreamLegacy, m->exit_ev, 0));
  }
  return 0;
}

constexpr size_t PIN_X = 256, PIN_OUT = 768, PIN_OUT_DOUBLES = 8 * (3 + 3 * 64), PIN_TOTAL = PIN_OUT + PIN_OUT_DOUBLES;

// Residual check of fits that used the int8 engine: tolerance on the componentwise backward error (0 switches the check off).  A
// backward-stable fp64 solve sits at a few eps sqrt(N) (~1e-14); 2e-13 (~1000 eps) is what keeps every result inside the
// cond(Ky) * eps allowances of the parity tests.
static double ozaki_check_tol() {
  static const double tol = [] {
    const char *e = getenv("GPB_OZAKI_CHECK_TOL");
    return e ? atof(e) : 2e-13;
  }();
  return tol;
}
static std::atomic<long long> g_engine_fallbacks{0};

static int g_overlap_min_n = 512;  // 0 disables the two-stream schedule (gpb_set_overlap)
static int g_config_epoch = 0;     // bumped by the tuning entry points: captured launch sequences are stale afterwards

static size_t carve(int n_cap, int d, int p, int cb, gpb_model *m, char *base) {
  const size_t np = round_up(std::max(n_cap, 1), TILE);
  const size_t nb = np / TILE;
  size_t off = 0;
  auto take = [&](size_t count, double **dst) {
    if (m && dst) *dst = reinterpret_cast<double *>(base + off);
    off += align256(count * sizeof(double));
  };
  double *fa = nullptr, *fm = nullptr, *fw = nullptr, *fpart = nullptr;
  take(np * np, &fa);
  take(np * np, &fm);
  take(np * np, &fw);
  take(std::max<size_t>(nb, 128) * np, &fpart);  // GEMV partials: nb * np (factor_solve), 8 * 16 * np (skinny products), 128 np (fused skinny call)
  take(skinny_fused_part3_doubles(d), m ? &m->sk_part3 : nullptr);
  take(kgrad_part_doubles((int)np, (int)np, d, 1), m ? &m->gpart : nullptr);
  take(np * (size_t)d, m ? &m->X : nullptr);
  take(np * (size_t)d, m ? &m->XsT : nullptr);
  take(np * (size_t)p, m ? &m->Yc : nullptr);
  take(np * (size_t)p, m ? &m->alpha : nullptr);
  take(np * (size_t)p, m ? &m->z : nullptr);
  take(d, m ? &m->ls_dev : nullptr);
  take(d, m ? &m->inv_ls_dev : nullptr);
  take(np * (size_t)d, m ? &m->XgT : nullptr);
  take((size_t)cb * d, m ? &m->XcgT : nullptr);
  take(d, m ? &m->gflag_dev : nullptr);
  take(d, m ? &m->ginv_dev : nullptr);
  take(d + 16, m ? &m->scal : nullptr);
  take(8, m ? &m->theta_dev : nullptr);
  take((size_t)cb * d, m ? &m->Xc : nullptr);
  take((size_t)cb * d, m ? &m->XcT : nullptr);
  take((size_t)cb * np, m ? &m->KxT : nullptr);
  take((size_t)cb * np, m ? &m->Vt : nullptr);
  take((size_t)cb * np, m ? &m->Ut : nullptr);
  take((size_t)cb * p, m ? &m->mu : nullptr);
  take(cb, m ? &m->var : nullptr);
  take((size_t)cb * d, m ? &m->dmu : nullptr);
  take((size_t)cb * d, m ? &m->dvar : nullptr);
  take(cb, m ? &m->fbuf : nullptr);
  take((size_t)cb * d, m ? &m->dfbuf : nullptr);
  take(cb, m ? &m->sdbuf : nullptr);
  take((size_t)cb * d, m ? &m->dsbuf : nullptr);
  take(64, m ? &m->topv : nullptr);
  double *ti = nullptr;
  take(64, &ti);
  double *finfo = nullptr;
  take(4, &finfo);
  if (m) {
    m->f.A = fa;
    m->f.Mi = fm;
    m->f.W = fw;
    m->f.part = fpart;
    m->topi = reinterpret_cast<long long *>(ti);
    m->f.info = reinterpret_cast<int *>(finfo);
  }
  return off;
}

extern "C" {

int gpb_version(void) { return 100; }
const char *gpb_last_error(void) { return g_err; }
long long gpb_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
int gpb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

size_t gpb_model_workspace_bytes(int n_cap, int d, int p, int cand_block) {
  if (n_cap < 1 || d < 1 || p < 1 || cand_block < 1) return 0;
  return carve(n_cap, d, p, round_up(cand_block, TILE), nullptr, nullptr);
}

int gpb_model_create(gpb_model **out, int kind, int ard, int d, int p, int n_cap, int cand_block, void *workspace,
                     size_t workspace_bytes, void *stream) {
  GPB_RANGE("gpb_model_create");
  GPB_REQUIRE(out != nullptr, "model_create: out is NULL");
  GPB_REQUIRE(kind == GPB_KERN_RBF || kind == GPB_KERN_MATERN52, "model_create: unknown kernel kind %d", kind);
  // 64 = what the predictive kernels (skinny moments, gradients_X) are instantiated for: a model that could be fitted but not queried
  // would be useless to the BO loop
  GPB_REQUIRE(d >= 1 && d <= 64, "model_create: input_dim %d out of range [1, 64]", d);
  GPB_REQUIRE(p >= 1 && p <= 16, "model_create: output_dim %d out of range [1, 16]", p);
  GPB_REQUIRE(n_cap >= 1 && cand_block >= 1, "model_create: capacities must be positive");
  GPB_REQUIRE(gpb_device_count() > 0, "model_create: no CUDA device visible -- libgpb200 has no CPU fallback");
  gpb_model *m = new gpb_model();
  // every failure path below releases what has been created so far
  auto fail = [&](int rc) {
    if (m->ov) factor_overlap_destroy(m->ov);
    if (m->pinned) cudaFreeHost(m->pinned);
    if (m->own_ws && m->ws) cudaFree(m->ws);
    if (m->entry_ev) cudaEventDestroy(m->entry_ev);
    if (m->exit_ev) cudaEventDestroy(m->exit_ev);
    if (m->own_stream) cudaStreamDestroy(m->stream);
    delete m;
    return rc;
  };
  m->kind = kind;
  m->ard = ard ? 1 : 0;
  m->d = d;
  m->p = p;
  m->n_cap = n_cap;
  m->np_cap = round_up(n_cap, TILE);
  m->cb = round_up(cand_block, TILE);
  m->nls = ard ? d : 1;
  m->ls.assign(m->nls, 1.0);
  m->stream = reinterpret_cast<cudaStream_t>(stream);
  if (m->stream == nullptr) {
    // The legacy default stream cannot be captured into a CUDA graph, and a BLOCKING stream must not be captured while another
    // thread uses the legacy stream (cudaErrorStreamCaptureImplicit).  The model gets a non-blocking stream of its own and orders
    // it against the legacy stream explicitly (check_device).
    if (cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking) == cudaSuccess) {
      m->own_stream = true;
      if (cudaEventCreateWithFlags(&m->entry_ev, cudaEventDisableTiming) != cudaSuccess ||
          cudaEventCreateWithFlags(&m->exit_ev, cudaEventDisableTiming) != cudaSuccess) {
        set_error("model_create: cudaEventCreate failed");
        return fail(-1);
      }
    } else {
      (void)cudaGetLastError();
      m->stream = nullptr;
      m->graph_failed = true;      // no private stream: stay on the legacy stream, never capture
    }
  }
  if (cudaGetDevice(&m->device) != cudaSuccess) {
    set_error("model_create: cudaGetDevice failed");
    return fail(-1);
  }
  const size_t need = carve(n_cap, d, p, m->cb, nullptr, nullptr);
  if (workspace) {
    if (workspace_bytes < need || (reinterpret_cast<uintptr_t>(workspace) & 255) != 0) {
      set_error("model_create: workspace too small or not 256-byte aligned (%zu < %zu)", workspace_bytes, need);
      return fail(-2);
    }
    m->ws = workspace;
  } else {
    cudaError_t e = cudaMalloc(&m->ws, need);
    if (e != cudaSuccess) {
      m->ws = nullptr;
      set_error("model_create: cudaMalloc(%zu) -> %s", need, cudaGetErrorString(e));
      return fail(-1);
    }
    m->own_ws = true;
  }
  carve(n_cap, d, p, m->cb, m, reinterpret_cast<char *>(m->ws));
  m->f.stream = m->stream;
  // arrival counter of the fused skinny call: zero once, the kernel re-arms it
  if (cudaMemsetAsync(m->sk_part3 + skinny_fused_part3_doubles(d) - 8, 0, 8 * sizeof(double), m->stream) != cudaSuccess) {
    set_error("model_create: cudaMemsetAsync failed");
    return fail(-1);
  }
  if (cudaMallocHost(&m->pinned, PIN_TOTAL * sizeof(double)) != cudaSuccess) {
    m->pinned = nullptr;
    set_error("model_create: cudaMallocHost failed");
    return fail(-1);
  }
  if (factor_overlap_create(&m->ov) != 0) {
    m->ov = nullptr;
    return fail(-1);
  }
  *out = m;
  return 0;
}

int gpb_model_destroy(gpb_model *m) {
  if (!m) return 0;
  cudaStreamSynchronize(m->stream);
  for (auto &g : m->graphs)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  // the int8 engine keeps a plane workspace per stream it was launched on: this model's streams are about to disappear
  if (m->own_stream) ozaki_release_stream(m->stream);
  if (m->ov) {
    ozaki_release_stream(m->ov->main);
    for (int d = 0; d < FactorOverlap::MAX_DEPTH; ++d) ozaki_release_stream(m->ov->side[d]);
  }
  factor_overlap_destroy(m->ov);
  if (m->lane_stream) {
    ozaki_release_stream(m->lane_stream);
    cudaStreamSynchronize(m->lane_stream);
    cudaStreamDestroy(m->lane_stream);
  }
  for (auto &e : m->lane_ev)
    if (e) cudaEventDestroy(e);
  if (m->lane_mem) cudaFree(m->lane_mem);
  if (m->entry_ev) cudaEventDestroy(m->entry_ev);
  if (m->exit_ev) cudaEventDestroy(m->exit_ev);
  if (m->own_stream) cudaStreamDestroy(m->stream);
  if (m->lp_buf) cudaFree(m->lp_buf);
  if (m->own_ws && m->ws) cudaFree(m->ws);
  if (m->pinned) cudaFreeHost(m->pinned);
  delete m;
  return 0;
}

int gpb_model_set_data(gpb_model *m, int n, const double *X, const double *Y, int dev) {
  GPB_RANGE("gpb_model_set_data");
  GPB_REQUIRE(m && X && Y, "set_data: NULL argument");
  GPB_TRY(check_device(m, "set_data"));
  GPB_REQUIRE(n >= 1 && n <= m->n_cap, "set_data: n = %d exceeds the model capacity %d", n, m->n_cap);
  m->n = n;
  m->np = round_up(n, TILE);
  m->f.n = n;
  m->f.np = m->np;
  const cudaMemcpyKind kind = dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  GPB_CUDA(cudaMemcpyAsync(m->X, X, (size_t)n * m->d * sizeof(double), kind, m->stream));
  // Y -> column blocks; stage the row-major copy in z (same size class) first when it comes from the host
  // (no temporary allocation: cudaMalloc / cudaFree per update cost milliseconds in a BO loop that appends a point per step)
  const double *Yd = Y;
  if (!dev) {
    GPB_CUDA(cudaMemcpyAsync(m->z, Y, (size_t)n * m->p * sizeof(double), cudaMemcpyHostToDevice, m->stream));
    Yd = m->z;
  }
  pack_cols_kernel<<<(m->p * m->np + 255) / 256, 256, 0, m->stream>>>(Yd, n, m->p, m->np, m->Yc);
  count_launch();
  GPB_CHECK_LAUNCH();
  GPB_CUDA(cudaStreamSynchronize(m->stream));
  m->have_data = true;
  m->scaled_valid = false;
  m->fitted = false;
  m->have_wi = false;
  m->wi_from = 0;
  return 0;
}

int gpb_model_set_gower(gpb_model *m, int enable, const int *discrete, const double *ranges) {
  GPB_REQUIRE(m, "set_gower: NULL model");
  GPB_TRY(check_device(m, "set_gower"));
  m->scaled_valid = false;
  m->fitted = false;
  m->have_wi = false;
  m->wi_from = 0;
  if (!enable) {
    m->gower = false;
    return 0;
  }
  GPB_REQUIRE(discrete && ranges, "set_gower: NULL argument");
  const int d = m->d;
  m->g_flag.assign(d, 0.0);
  m->g_inv.assign(d, 1.0);
  for (int q = 0; q < d; ++q) {
    if (discrete[q]) {
      m->g_flag[q] = 1.0;
    } else {
      GPB_REQUIRE(ranges[q] > 0 && std::isfinite(ranges[q]), "set_gower: range of continuous dimension %d must be positive", q);
      m->g_inv[q] = ranges[q];   // scale_transpose divides by its scale vector
    }
  }
  GPB_CUDA(cudaMemcpyAsync(m->gflag_dev, m->g_flag.data(), d * sizeof(double), cudaMemcpyHostToDevice, m->stream));
  GPB_CUDA(cudaMemcpyAsync(m->ginv_dev, m->g_inv.data(), d * sizeof(double), cudaMemcpyHostToDevice, m->stream));
  GPB_CUDA(cudaStreamSynchronize(m->stream));
  m->gower = true;
  return 0;
}

int gpb_model_set_theta(gpb_model *m, double variance, const double *lengthscale, double noise) {
  GPB_RANGE("gpb_model_set_theta");
  GPB_REQUIRE(m && lengthscale, "set_theta: NULL argument");
  {
    // bit-identical hyper-parameters keep the factorisation (what makes gpb_model_append usable after GP.set_XY)
    bool same = m->variance == variance && m->noise == noise;
    for (int q = 0; q < m->nls; ++q) same = same && m->ls[q] == lengthscale[q];
    if (same) return 0;
  }
  m->variance = variance;
  m->noise = noise;
  for (int q = 0; q < m->nls; ++q) m->ls[q] = lengthscale[q];
  m->scaled_valid = false;
  m->fitted = false;
  m->have_wi = false;
  m->wi_from = 0;
  return 0;
}

static int ensure_scaled(gpb_model *m) {
  if (m->scaled_valid) return 0;
  GPB_TRY(set_ls(m->ls.data(), m->nls, m->d, m->ls_dev, m->inv_ls_dev, m->stream));
  GPB_TRY(launch_scale_transpose(m->X, m->n, m->d, m->ls_dev, m->XsT, m->np, m->stream));
  if (m->gower) GPB_TRY(launch_scale_transpose(m->X, m->n, m->d, m->ginv_dev, m->XgT, m->np, m->stream, 1));
  m->scaled_valid = true;
  return 0;
}

// Coordinates, flags and variance factor the covariance kernel reads: Euclidean scaled (default) or the Gower set.
struct KCoords {
  const double *XT, *gflag;
  double var;
};
static KCoords train_coords(const gpb_model *m) {
  if (!m->gower) return {m->XsT, nullptr, m->variance};
  return {m->XgT, m->gflag_dev, std::pow(m->variance, m->d)};
}

static int ensure_wi(gpb_model *m) {
  if (m->have_wi) return 0;
  GPB_TRY(m->wi_from > 0 ? factor_potri_append(m->f, m->wi_from) : factor_potri(m->f));
  m->have_wi = true;
  m->wi_from = 0;
  return 0;
}

// append_from > 0: the leading append_from x append_from block of the factorisation is valid for the current hyper-parameters
// (gpb_model_append); only the block rows from there on are built and factorised.
// The launch sequence of one evaluation through the general path, results to the pinned block; no synchronisation, so that it
// can be captured.  theta != NULL (graph capture): hyper-parameters come from device memory fed by the pinned block
// [136] variance, [137] diag_add, [138 ..) ls[d], inv_ls[d].
static int fit_launch_general(gpb_model *m, int want_grad, double extra_jitter, int append_from, const double *theta) {
  const int n = m->n, np = m->np, d = m->d, p = m->p;
  if (theta) {
    GPB_CUDA(cudaMemcpyAsync(m->theta_dev, m->pinned + 136, 2 * sizeof(double), cudaMemcpyHostToDevice, m->stream));
    GPB_CUDA(cudaMemcpyAsync(m->ls_dev, m->pinned + 138, d * sizeof(double), cudaMemcpyHostToDevice, m->stream));
    GPB_CUDA(cudaMemcpyAsync(m->inv_ls_dev, m->pinned + 138 + d, d * sizeof(double), cudaMemcpyHostToDevice, m->stream));
    GPB_TRY(launch_scale_transpose(m->X, n, d, m->ls_dev, m->XsT, np, m->stream));
    m->scaled_valid = true;
  } else {
    GPB_TRY(ensure_scaled(m));
  }
  // Ky = K + (noise + 1e-8 [+ jitter]) I     exact_gaussian_inference.py:55-56
  const KCoords kc = train_coords(m);
  GPB_TRY(launch_kmat(m->kind, kc.XT, np, kc.XT, np, d, n, n, kc.var, m->noise + 1e-8 + extra_jitter, 3, m->f.A, np, np, np,
                      m->stream, kc.gflag, append_from, theta));
  // fork threshold: the top three levels of the recursion (more forks only add cross-stream latency, scripts/overlap_sweep.py)
  static const int fork_div = std::max(1, env_int("GPB_FORK_DIV", 8));
  // up to 4096 rows every level forks (each T12 taken off the critical path is one 6 - 15 us launch less on it: 3.750 -> 3.726 ms at
  // N = 4096, 0.441 -> 0.424 at N = 1024); above, only the top three levels (deeper forks cost 0.1 ms at N = 16384), scripts/overlap_min_probe.py
  const int fork_min_n = g_overlap_min_n > 0 ? (np <= 4096 && g_overlap_min_n == 512 ? 128 : std::max(g_overlap_min_n, np / fork_div)) : 0;
  if (m->ov && fork_min_n > 0 && np >= fork_min_n) {
    // critical path on the model's high-priority stream, T21 products on its low-priority side streams (gpb_chol.cu)
    Factor fo = m->f;
    fo.stream = m->ov->main;
    fo.ov = m->ov;
    m->ov->min_n = fork_min_n;
    GPB_CUDA(cudaEventRecord(m->ov->enter, m->stream));
    GPB_CUDA(cudaStreamWaitEvent(m->ov->main, m->ov->enter, 0));
    const int rc = append_from > 0 ? factor_append(fo, append_from) : factor_potrf_inv(fo);
    m->f.l_pending = fo.l_pending;
    m->f.l_from = fo.l_from;
    GPB_CUDA(cudaEventRecord(m->ov->leave, m->ov->main));
    GPB_CUDA(cudaStreamWaitEvent(m->stream, m->ov->leave, 0));
    GPB_TRY(rc);
  } else {
    GPB_TRY(append_from > 0 ? factor_append(m->f, append_from) : factor_potrf_inv(m->f));
  }
  m->engine_used = ozaki_min_n() > 0 && np >= ozaki_min_n() && append_from == 0;
  const bool check = m->engine_used && p == 1 && ozaki_check_tol() > 0.0;
  // alpha = M^T (M y), log det and alpha . y only read M and the diagonal of L, like Ky^-1 = M^T M: with gradients wanted they run on a
  // side stream underneath that product instead of in front of it (76 us of bandwidth-bound passes at N = 4096, 0.4 ms at 16384)
  static const bool side_on = env_int("GPB_SIDE_SOLVE", 1) != 0;
  const bool side = side_on && want_grad && m->ov != nullptr && !check && !m->have_wi;
  constexpr int SD = FactorOverlap::MAX_DEPTH - 1;        // the deepest side stream: the recursion never forks that far down
  Factor fs = m->f;
  cudaStream_t ss = m->stream;
  if (side) {
    ss = m->ov->side[SD];
    fs.stream = ss;
    fs.ov = nullptr;
    GPB_CUDA(cudaEventRecord(m->ov->fork[SD], m->stream));
    GPB_CUDA(cudaStreamWaitEvent(ss, m->ov->fork[SD], 0));
  }
  GPB_TRY(factor_solve(fs, m->Yc, p, m->z, m->alpha));
  // The int8 engine's products are accurate relative to the largest entry of an operand ROW, not entry by entry: when
  // it took part in this factorisation, measure what that did to the solve -- the componentwise backward error of Ky alpha = y,
  // against a Ky rebuilt from the inputs (W is free between the factorisation and the inverse) -- and let fit_core fall back.
  if (check) {
    GPB_TRY(factor_finalize_L(m->f));
    GPB_TRY(launch_kmat(m->kind, kc.XT, np, kc.XT, np, d, n, n, kc.var, m->noise + 1e-8 + extra_jitter, 1, m->f.W, np, np, np, m->stream,
                        kc.gflag, 0, theta));
    GPB_TRY(launch_solve_residual(m->f.W, np, n, m->alpha, m->Yc, m->scal + d + 8, m->stream));
  }
  GPB_TRY(factor_logdet(fs, m->scal + 0));
  dot_kernel<<<1, 1024, 0, ss>>>(m->alpha, m->Yc, p * np, m->scal + 1);
  count_launch();
  GPB_CHECK_LAUNCH();
  if (side) GPB_CUDA(cudaEventRecord(m->ov->join[SD], ss));
  if (want_grad) {
    GPB_TRY(ensure_wi(m));
    if (side) GPB_CUDA(cudaStreamWaitEvent(m->stream, m->ov->join[SD], 0));
    GPB_TRY(launch_kgrad(m->kind, 1, m->XsT, np, m->XsT, np, d, n, n, m->variance, m->f.W, np, m->alpha, np, p, m->gpart,
                         m->scal + 2, m->stream, theta));
    // Gower patch: only the variance term sees the patched K (stationary.py:224); it overwrites kgrad's Euclidean one
    if (m->gower)
      GPB_TRY(launch_kvar_gower(m->kind, 1, kc.XT, np, kc.XT, np, d, n, n, kc.var, kc.gflag, m->f.W, np, m->alpha, np, p, m->gpart,
                                m->scal + 2, m->stream));
  }
  GPB_CUDA(cudaMemcpyAsync(m->pinned, m->scal, (d + 9) * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
  GPB_CUDA(cudaMemcpyAsync(m->pinned + 128, m->f.info, sizeof(int), cudaMemcpyDeviceToHost, m->stream));
  return 0;
}

// Replay (or first capture) of the general launch sequence as a CUDA graph.  Returns 0 when the evaluation has been enqueued,
// 1 when the caller must take the ordinary path (first sight of this shape, or graphs unavailable), < 0 on error.
static int fit_launch_graph(gpb_model *m, int want_grad, double extra_jitter) {
  gpb_model::FitGraph &g = m->graphs[want_grad ? 1 : 0];
  const int n = m->n, np = m->np, d = m->d;
  if (g.n != n || g.np != np || g.epoch != g_config_epoch) {   // new shape: one ordinary evaluation first (it also configures
    if (g.exec) cudaGraphExecDestroy(g.exec);                  // the kernels' attributes, which must not happen inside a capture)
    g = gpb_model::FitGraph();
    g.n = n;
    g.np = np;
    g.epoch = g_config_epoch;
    g.seen = 1;
    return 1;
  }
  m->pinned[136] = m->variance;
  m->pinned[137] = m->noise + 1e-8 + extra_jitter;
  for (int q = 0; q < d; ++q) {
    const double l = m->ls[m->nls == 1 ? 0 : q];
    m->pinned[138 + q] = l;
    m->pinned[138 + d + q] = 1.0 / l;
  }
  if (!g.exec) {
    const long long l0 = g_launches.load(std::memory_order_relaxed);
    if (cudaStreamBeginCapture(m->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
      (void)cudaGetLastError();
      m->graph_failed = true;
      return 1;
    }
    const int rc = fit_launch_general(m, want_grad, extra_jitter, 0, m->theta_dev);
    cudaGraph_t graph = nullptr;
    const cudaError_t e = cudaStreamEndCapture(m->stream, &graph);
    if (rc != 0 || e != cudaSuccess || graph == nullptr || cudaGraphInstantiate(&g.exec, graph, 0) != cudaSuccess) {
      (void)cudaGetLastError();
      if (graph) cudaGraphDestroy(graph);
      g.exec = nullptr;
      m->graph_failed = true;   // this model stays on the ordinary path
      m->scaled_valid = false;
      m->have_wi = false;
      return 1;
    }
    cudaGraphDestroy(graph);
    g.launches = g_launches.load(std::memory_order_relaxed) - l0;
    g.l_pending = m->f.l_pending;
    g.l_from = m->f.l_from;
    g.have_wi = m->have_wi;
  } else {
    count_launch((int)g.launches);
  }
  GPB_CUDA(cudaGraphLaunch(g.exec, m->stream));
  m->scaled_valid = true;
  m->f.l_pending = g.l_pending;
  m->f.l_from = g.l_from;
  m->have_wi = g.have_wi;
  m->wi_from = 0;
  return 0;
}

static int fit_core(gpb_model *m, int want_grad, double extra_jitter, double *out, int append_from) {
  // Hyper-parameters outside the domain (an L-BFGS-B line search can push the transformed parameters to 0, inf or NaN):
  // the reference's NumPy path turns those into NaNs and ends in jitchol's LinAlgError (linalg.py:62-75), which paramz
  // catches.  Same contract here: GPB_ERR_DOMAIN -> LinAlgError in the Python layer.  variance == 0 is legal (K = 0).
  {
    bool ok = std::isfinite(m->variance) && m->variance >= 0 && std::isfinite(m->noise) && m->noise >= 0 &&
              std::isfinite(extra_jitter);
    for (int q = 0; q < m->nls; ++q) ok = ok && std::isfinite(m->ls[q]) && m->ls[q] > 0 && std::isfinite(1.0 / m->ls[q]);
    if (!ok) {
      set_error("fit: hyper-parameters outside the domain (variance >= 0, lengthscale > 0, noise >= 0, all finite)");
      return GPB_ERR_DOMAIN;
    }
  }
  m->fitted = false;
  m->have_wi = false;
  m->jitter = extra_jitter;
  const int n = m->n, np = m->np, d = m->d, p = m->p;
  static const int tiny_on = env_int("GPB_TINY", 1) != 0;
  const bool tiny = tiny_on && np == TILE && !m->gower && p == 1 && d <= 32 && append_from == 0;
  if (tiny) {
    // N <= 128 (an ordinary BO run): the whole evaluation -- input scaling included -- in one kernel instead of a dozen launches
    // and two small uploads; the status travels with the results
    double lsq[32];
    for (int q = 0; q < d; ++q) lsq[q] = m->ls[m->nls == 1 ? 0 : q];
    GPB_TRY(launch_tiny_fit(m->kind, m->X, lsq, m->XsT, m->ls_dev, m->inv_ls_dev, n, d, m->variance, m->noise + 1e-8 + extra_jitter, m->Yc,
                            m->f, m->z, m->alpha, m->scal, want_grad));
    m->scaled_valid = true;
    if (want_grad) m->have_wi = true;
  } else {
    // replay pays up to here (scripts/small_n_perf.py: -16% at N = 256, -6% at 1024, nothing at 4096)
    static const int graph_max_np = env_int("GPB_GRAPH_MAX_NP", 2048);
    int rc = 1;
    if (np <= graph_max_np && append_from == 0 && !m->gower && d <= 32 && !m->graph_failed && !gemm_profile_is_on() &&
        !(ozaki_min_n() > 0 && np >= ozaki_min_n()))   // the int8 engine allocates its digit workspace on first use: not capturable
      rc = fit_launch_graph(m, want_grad, extra_jitter);
    if (rc < 0) return rc;
    if (rc == 1) GPB_TRY(fit_launch_general(m, want_grad, extra_jitter, append_from, nullptr));
  }
  if (tiny) GPB_CUDA(cudaMemcpyAsync(m->pinned, m->scal, (d + 5) * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
  GPB_CUDA(cudaStreamSynchronize(m->stream));
  const int info = tiny ? (int)m->pinned[d + 4] : *reinterpret_cast<int *>(m->pinned + 128);
  if (info != 0) {
    set_error("fit: matrix not positive definite (leading minor %d)", info);
    return info > n ? n : info;
  }
  m->engine_ok = true;
  m->last_residual = -1.0;
  if (!tiny && m->engine_used && p == 1 && ozaki_check_tol() > 0.0) {
    m->last_residual = m->pinned[d + 8];
    if (!(m->last_residual <= ozaki_check_tol())) {
      // the engine's row-wise accuracy was not enough for this matrix: the same evaluation on the fp64 DMMA engine
      g_engine_fallbacks.fetch_add(1, std::memory_order_relaxed);
      ozaki_suppress(1);
      m->have_wi = false;
      const int rc2 = fit_launch_general(m, want_grad, extra_jitter, append_from, nullptr);
      const cudaError_t e2 = rc2 == 0 ? cudaStreamSynchronize(m->stream) : cudaSuccess;
      ozaki_suppress(0);
      GPB_TRY(rc2);
      GPB_CUDA(e2);
      m->engine_ok = false;
      m->engine_used = false;
      const int info2 = *reinterpret_cast<int *>(m->pinned + 128);
      if (info2 != 0) {
        set_error("fit: matrix not positive definite (leading minor %d)", info2);
        return info2 > n ? n : info2;
      }
    }
  }
  const double logdet = m->pinned[0], ay = m->pinned[1];
  out[0] = 0.5 * (-(double)n * p * log(2.0 * M_PI) - (double)p * logdet - ay);  // exact_gaussian_inference.py:62
  if (want_grad) {
    const double *g = m->pinned + 2;
    out[1] = g[0] / m->variance;                          // stationary.py:224
    if (m->ard) {
      for (int q = 0; q < d; ++q) out[2 + q] = -g[2 + q] / m->ls[q];   // stationary.py:233,268 in scaled coordinates
    } else {
      double sacc = 0.0;
      for (int q = 0; q < d; ++q) sacc += g[2 + q];
      out[2] = -sacc / m->ls[0];                         // stationary.py:237-238
    }
    out[2 + m->nls] = g[1];                               // exact_gaussian_inference.py:72, gaussian.py:78-79
  }
  m->fitted = true;
  m->l_valid = true;
  return 0;
}

int gpb_model_fit(gpb_model *m, int want_grad, double extra_jitter, double *out) {
  GPB_RANGE("gpb_model_fit");
  GPB_REQUIRE(m && out, "fit: NULL argument");
  GPB_TRY(check_device(m, "fit"));
  GPB_REQUIRE(m->have_data, "fit: set_data has not been called");
  m->wi_from = 0;
  return fit_core(m, want_grad, extra_jitter, out, 0);
}

int gpb_model_append(gpb_model *m, int b, const double *Xnew, const double *Yall, int dev, int want_grad, double *out) {
  GPB_RANGE("gpb_model_append");
  GPB_REQUIRE(m && Xnew && Yall && out, "append: NULL argument");
  GPB_TRY(check_device(m, "append"));
  GPB_REQUIRE(m->fitted && m->jitter == 0.0, "append: needs a model fitted (without extra jitter) for the current hyper-parameters");
  GPB_REQUIRE(m->l_valid, "append: the model adopted a broadcast state without the factor L (gpb_model_adopt_state have_mask bit 0)");
  GPB_REQUIRE(b >= 1 && m->n + b <= m->n_cap, "append: %d + %d points exceed the model capacity %d", m->n, b, m->n_cap);
  const int n_old = m->n, np_old = m->np, n_new = n_old + b, np_new = round_up(n_new, TILE), d = m->d;
  cudaStream_t s = m->stream;
  AllocStream alloc_scope(s);
  // block rows from h on are rebuilt: the last, partially filled block row of the old factor and everything new
  const int h = (n_old / TILE) * TILE;
  // Ky^-1: keep the leading h block across the append (factor_potri_append completes it in O(N^2 r))
  int wi_keep = 0;
  if (h > 0 && m->have_wi) {
    GPB_TRY(factor_potri_downdate(m->f, h, np_old));
    wi_keep = h;
  } else if (h > 0 && m->wi_from > 0) {
    wi_keep = m->wi_from;                               // an earlier append whose inverse was never asked for
  }
  GPB_TRY(factor_finalize_L(m->f));                     // moves L out of W (no-op when W holds the inverse)
  if (np_new != np_old) {
    // the three N x N matrices are stored with leading dimension np: re-stride them through a temporary (or, when that
    // allocation fails, through W, giving up the inverse)
    double *tmp = nullptr;
    const size_t cnt = (size_t)np_old * np_old;
    if (cudaMalloc(&tmp, cnt * sizeof(double)) != cudaSuccess) {   // (multi-GB: plain cudaMalloc is 3-5x faster than growing the async pool)
      (void)cudaGetLastError();
      tmp = m->f.W;
      wi_keep = 0;
    }
    for (double *mat : {m->f.A, m->f.Mi, m->f.W}) {
      if (mat == m->f.W && (wi_keep == 0 || tmp == m->f.W)) continue;
      GPB_CUDA(cudaMemcpyAsync(tmp, mat, cnt * sizeof(double), cudaMemcpyDeviceToDevice, s));
      GPB_TRY(launch_copy2d(mat, np_new, tmp, np_old, np_old, np_old, s));
    }
    if (tmp != m->f.W) {
      GPB_CUDA(cudaStreamSynchronize(s));
      GPB_CUDA(cudaFree(tmp));
    }
  }
  GPB_CUDA(cudaMemcpyAsync(m->X + (size_t)n_old * d, Xnew, (size_t)b * d * sizeof(double),
                           dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s));
  m->n = n_new;
  m->np = np_new;
  m->f.n = n_new;
  m->f.np = np_new;
  const double *Yd = Yall;
  if (!dev) {
    GPB_CUDA(cudaMemcpyAsync(m->z, Yall, (size_t)n_new * m->p * sizeof(double), cudaMemcpyHostToDevice, s));
    Yd = m->z;
  }
  pack_cols_kernel<<<(m->p * np_new + 255) / 256, 256, 0, s>>>(Yd, n_new, m->p, np_new, m->Yc);
  count_launch();
  GPB_CHECK_LAUNCH();
  GPB_CUDA(cudaStreamSynchronize(s));
  m->scaled_valid = false;
  m->wi_from = wi_keep;
  const int rc = fit_core(m, want_grad, 0.0, out, h);
  if (rc != 0) {                    // the factor is in an undefined state: the caller must run a full fit
    m->fitted = false;
    m->wi_from = 0;
  }
  return rc;
}

int gpb_model_get(gpb_model *m, const char *what, double *dst, int ld, int dev) {
  GPB_RANGE("gpb_model_get");
  GPB_REQUIRE(m && what && dst, "get: NULL argument");
  GPB_TRY(check_device(m, "get"));
  const std::string w(what);
  const int n = m->n, np = m->np, p = m->p;
  cudaStream_t s = m->stream;
  AllocStream alloc_scope(s);
  const bool is_mat = (w == "L" || w == "Li" || w == "Wi" || w == "K" || w == "dL_dK");
  GPB_REQUIRE(is_mat || w == "alpha", "get: unknown quantity '%s'", what);
  if (w != "K") GPB_REQUIRE(m->fitted, "get: model has not been fitted");
  const int cols = is_mat ? n : p;
  GPB_REQUIRE(ld >= cols, "get: ld too small");
  DevBuf tmp;
  double *ddst = dst;
  int ldd = ld;
  if (!dev) {
    GPB_TRY(tmp.alloc((size_t)n * cols * sizeof(double)));
    ddst = tmp.d();
    ldd = cols;
  }
  const dim3 grid((n + 255) / 256, n), gsym((n + 31) / 32, (n + 31) / 32);
  if (w == "L") {
    GPB_REQUIRE(m->l_valid, "get: the model adopted a broadcast state without the factor L");
    GPB_TRY(factor_finalize_L(m->f));
    tril_copy_kernel<<<grid, 256, 0, s>>>(m->f.A, np, ddst, ldd, n);
  } else if (w == "Li") {
    tril_copy_kernel<<<grid, 256, 0, s>>>(m->f.Mi, np, ddst, ldd, n);
  } else if (w == "Wi") {
    GPB_TRY(ensure_wi(m));
    sym_copy_kernel<<<gsym, dim3(32, 8), 0, s>>>(m->f.W, np, ddst, ldd, n);
  } else if (w == "dL_dK") {
    GPB_TRY(ensure_wi(m));
    dl_dk_kernel<<<grid, 256, 0, s>>>(m->f.W, np, m->alpha, np, p, n, ddst, ldd);
  } else if (w == "K") {
    GPB_REQUIRE(m->have_data, "get: no data");
    GPB_TRY(ensure_scaled(m));
    const KCoords kc = train_coords(m);
    GPB_TRY(launch_kmat(m->kind, kc.XT, np, kc.XT, np, m->d, n, n, kc.var, 0.0, 0, ddst, ldd, np, np, s, kc.gflag));
  } else {
    unpack_cols_kernel<<<(n * p + 255) / 256, 256, 0, s>>>(m->alpha, n, p, np, ddst);
  }
  count_launch();
  GPB_CHECK_LAUNCH();
  if (!dev) {
    GPB_CUDA(cudaMemcpy2DAsync(dst, (size_t)ld * sizeof(double), ddst, (size_t)ldd * sizeof(double), (size_t)cols * sizeof(double),
                               n, cudaMemcpyDeviceToHost, s));
  }
  GPB_CUDA(cudaStreamSynchronize(s));
  return 0;
}

// ---- predictive pipeline over one candidate block -------------------------------------------------------------------
// level 0: mean only; 1: + variance; 2: + gradients (dmu, dvar); 3: mean and its gradient only (estimate_L)
// The two triangular products of a candidate block go through the int8 engine when it is switched on, the model is large enough
// and the block has enough rows to fill 256-row tiles.
// 8 digits here, not 7: the predictive variance sigma^2 - sum (M k*)^2 cancels, and the digit scheme is accurate relative to the
// largest entry of a row of M, not entry by entry (with 7 digits LCB values of an ill-conditioned RBF model moved by 4e-10).
// 8 digits, or 18 moduli (62 bits per operand) when the engine is configured in its modular mode
#define OZAKI_PREDICT_DIGITS ozaki_predict_planes()
static inline bool ozaki_predict(const gpb_model *m, int np, int cpad) {
  return m->engine_ok && ozaki_min_n() > 0 && np >= ozaki_min_n() && cpad >= 1024;
}

// allow_fused = false: the caller reads the block's intermediates afterwards (XcT, and Vt in its row-per-candidate layout -- the fused
// M <= 8 kernel keeps neither: it never forms XcT and parks k'/r and M k* in rows 0-7 / 8-15 of Vt)
static int predict_block(gpb_model *m, const double *Xc, int mcb, int dev, int level, int include_likelihood, bool allow_fused = true) {
  const int n = m->n, np = m->np, d = m->d, p = m->p;
  const int cpad = round_up(mcb, TILE);
  cudaStream_t s = m->stream;
  AllocStream alloc_scope(s);
  if (!dev && (size_t)mcb * d <= 512) {
    // a pageable source makes cudaMemcpyAsync stage and wait; the pinned block turns the small uploads of the M <= 8 calls into
    // plain asynchronous copies (every entry point synchronises before it returns, so the block is free again by the next call)
    memcpy(m->pinned + PIN_X, Xc, (size_t)mcb * d * sizeof(double));
    GPB_CUDA(cudaMemcpyAsync(m->Xc, m->pinned + PIN_X, (size_t)mcb * d * sizeof(double), cudaMemcpyHostToDevice, s));
  } else {
    GPB_CUDA(cudaMemcpyAsync(m->Xc, Xc, (size_t)mcb * d * sizeof(double), dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s));
  }
  // A handful of candidates (the M = 1 .. 8 calls of the L-BFGS-B refinements, optimizer.py:46-51, coalesced by the host's
  // LockstepEvaluator): everything from the covariance row to the four reductions in ONE persistent cooperative kernel
  // (gpb_skinny.cu) -- two streaming passes over the triangle of M.  Kx and U live in the first 16 rows of KxT, Dk and Z in the first 16 of Vt.
  static const int fused_on = env_int("GPB_SKINNY_FUSED", 1);
  if (allow_fused && mcb <= 8 && p == 1 && !m->gower && fused_on && level >= 1)
    return launch_skinny_fused(m->kind, m->f.Mi, np, n, d, mcb, level, m->XsT, m->Xc, m->ls_dev, m->inv_ls_dev, m->alpha, m->variance,
                               m->variance + (include_likelihood ? m->noise : 0.0), m->KxT, m->Vt, m->Vt + (size_t)8 * np, m->KxT + (size_t)8 * np, m->f.part, m->sk_part3,
                               m->mu, m->var, m->dmu, m->dvar, s);
  GPB_TRY(launch_scale_transpose(m->Xc, mcb, d, m->ls_dev, m->XcT, cpad, s));
  // KxT[c][n] = k(x*_c, x_n)                                        posterior.py:275 (stored transposed)
  const KCoords kc = train_coords(m);
  const double *XcK = m->XcT;
  if (m->gower) {
    GPB_TRY(launch_scale_transpose(m->Xc, mcb, d, m->ginv_dev, m->XcgT, cpad, s, 1));
    XcK = m->XcgT;
  }
  GPB_TRY(launch_kmat(m->kind, XcK, cpad, kc.XT, np, d, mcb, n, kc.var, 0.0, 2, m->KxT, np, cpad, np, s, kc.gflag));
  // A handful of candidates (the M = 1 calls of the L-BFGS-B refinement, optimizer.py:46-51): one bandwidth-bound pass over
  // the triangle of M per product instead of a 128-row padded GEMM, and the row reductions (mean, variance, both input
  // gradients) split over the training points in one fused pass (launch_skinny_moments) instead of one warp per candidate.
  const bool skinny = mcb <= 8;
  // the skinny products run on THIS block's stream: in the two-lane scoring pass that is not the factor's own stream (they used to be
  // issued there, unordered against the lane's covariance rows and reductions: found by the random sweep under GPB_SKINNY_FUSED=0;
  // the same route serves Gower models by default)
  Factor fl = m->f;
  fl.stream = s;
  fl.ov = nullptr;
  if (skinny && p == 1 && d <= 32) {      // (round-1 route: Gower kernel, GPB_SKINNY_FUSED=0; wider inputs take the generic kernels below)
    const double var_base = m->variance + (include_likelihood ? m->noise : 0.0);
    const double *Vt = nullptr, *Ut = nullptr;
    if (level == 1 || level == 2) {
      const int c = mcb <= 1 ? 1 : mcb <= 2 ? 2 : mcb <= 4 ? 4 : 8;   // rows mcb .. c of KxT are zero (mode 2 padding)
      GPB_TRY(factor_skinny_products(fl, c, m->KxT, np, m->Vt, np, level >= 2 ? m->Ut : nullptr, np, m->f.part));
      Vt = m->Vt;
      if (level == 2) Ut = m->Ut;
    }
    // mu = Kx^T alpha; var = Kdiag - sum(tmp^2) (+ noise); dmu = gradients_X(alpha^T, X*, X); dvar = gradients_X(-2 Kx^T Wi, X*, X)
    //                                                         posterior.py:276,294-295, gaussian.py:109, core/gp.py:431-434,450-453
    return launch_skinny_moments(m->kind, m->KxT, Vt, Ut, np, mcb, n, m->alpha, m->XcT, cpad, m->XsT, np, d, m->variance, m->inv_ls_dev,
                                 var_base, level >= 2 ? 1 : 0, m->f.part, m->mu, m->var, m->dmu, m->dvar, s);
  }
  // mu = Kx^T alpha                                                 posterior.py:276
  GPB_TRY(launch_rowdot(m->KxT, np, mcb, n, m->alpha, np, p, m->mu, s));
  if (level == 3) {
    GPB_REQUIRE(p == 1, "predictive gradients are implemented for a single output column (got %d)", p);
    GPB_TRY(launch_gradx(m->kind, m->XcT, cpad, mcb, m->XsT, np, n, d, m->variance, m->inv_ls_dev, m->alpha, 0, 1.0, 0, nullptr, 0,
                         0.0, m->dmu, nullptr, d, s));
    return 0;
  }
  if (level >= 1 && skinny) {
    const int c = mcb <= 1 ? 1 : mcb <= 2 ? 2 : mcb <= 4 ? 4 : 8;   // rows mcb .. c of KxT are zero (mode 2 padding)
    GPB_TRY(factor_skinny_products(fl, c, m->KxT, np, m->Vt, np, level >= 2 ? m->Ut : nullptr, np, m->f.part));
    GPB_TRY(launch_var_from_vt(m->Vt, np, mcb, n, m->variance + (include_likelihood ? m->noise : 0.0), m->var, s));
  } else if (level >= 1) {
    // Vt = KxT M^T  (== (L^-1 Kx)^T: dtrtrs of posterior.py:293 as a product with the explicit inverse factor)
    GemmArgs g{m->KxT, np, m->f.Mi, np, m->Vt, np, cpad, np, np, 1.0, 0.0, 0, 0, 2};
    // experimental int8 engine (gpb_ozaki.cu) for big candidate blocks; the digit planes of L^-1 are cut once per factorisation
    if (ozaki_predict(m, np, cpad)) GPB_TRY(ozaki_gemm_launch(LAYOUT_ROWK, LAYOUT_ROWK, g, 0, 1, OZAKI_PREDICT_DIGITS, s, 1));
    else GPB_TRY(gemm_launch(LAYOUT_ROWK, LAYOUT_ROWK, g, s));
    // var = Kdiag - sum(tmp^2) (+ noise)                            posterior.py:294-295, gaussian.py:109
    GPB_TRY(launch_var_from_vt(m->Vt, np, mcb, n, m->variance + (include_likelihood ? m->noise : 0.0), m->var, s));
  }
  if (level >= 2) {
    GPB_REQUIRE(p == 1, "predictive gradients are implemented for a single output column (got %d)", p);
    // Ut = Vt M = (Ky^-1 Kx)^T                                      core/gp.py:450-451 (woodbury_inv product as 2nd triangular product)
    if (!skinny) {
      GemmArgs g{m->Vt, np, m->f.Mi, np, m->Ut, np, cpad, np, np, 1.0, 0.0, 0, 2, 0};
      if (ozaki_predict(m, np, cpad)) GPB_TRY(ozaki_gemm_launch(LAYOUT_ROWK, LAYOUT_COLK, g, 0, 2, OZAKI_PREDICT_DIGITS, s, 2));
      else GPB_TRY(gemm_launch(LAYOUT_ROWK, LAYOUT_COLK, g, s));
    }
    // dmu = gradients_X(alpha^T, X*, X); dvar = gradients_X(-2 Kx^T Wi, X*, X)     core/gp.py:431-434,450-453
    GPB_TRY(launch_gradx(m->kind, m->XcT, cpad, mcb, m->XsT, np, n, d, m->variance, m->inv_ls_dev, m->alpha, 0, 1.0, 0, m->Ut, np,
                         -2.0, m->dmu, m->dvar, d, s));
  }
  return 0;
}

static int copy_out(double *dst, const double *src_dev, size_t count, int dev, cudaStream_t s) {
  if (!dst) return 0;
  GPB_CUDA(cudaMemcpyAsync(dst, src_dev, count * sizeof(double), dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, s));
  return 0;
}

// Small host results (the M <= 8 calls): device -> pinned block asynchronously, unpacked by finish_staged() after the call's one
// synchronisation, instead of one blocking pageable copy per output array.
static int copy_out_staged(gpb_model *m, double *dst, const double *src_dev, size_t count) {
  if (!dst) return 0;
  if (m->staged_used + count > PIN_OUT_DOUBLES) return copy_out(dst, src_dev, count, 0, m->stream);
  GPB_CUDA(cudaMemcpyAsync(m->pinned + PIN_OUT + m->staged_used, src_dev, count * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
  m->staged.push_back({dst, m->staged_used, count});
  m->staged_used += count;
  return 0;
}
static int finish_staged(gpb_model *m) {
  const cudaError_t e = cudaStreamSynchronize(m->stream);
  if (e == cudaSuccess)
    for (const auto &st : m->staged) memcpy(st.dst, m->pinned + PIN_OUT + st.off, st.count * sizeof(double));
  m->staged.clear();
  m->staged_used = 0;
  GPB_CUDA(e);
  return 0;
}

int gpb_model_predict(gpb_model *m, int mc, const double *Xc, int include_likelihood, double *mu, double *var, int dev) {
  GPB_RANGE("gpb_model_predict");
  GPB_REQUIRE(m && Xc, "predict: NULL argument");
  GPB_TRY(check_device(m, "predict"));
  GPB_REQUIRE(m->fitted, "predict: model has not been fitted");
  GPB_REQUIRE(mc >= 0, "predict: negative candidate count");
  for (int c0 = 0; c0 < mc; c0 += m->cb) {
    const int mcb = std::min(m->cb, mc - c0);
    GPB_TRY(predict_block(m, Xc + (size_t)c0 * m->d, mcb, dev, var ? 1 : 0, include_likelihood));
    GPB_TRY(copy_out(mu ? mu + (size_t)c0 * m->p : nullptr, m->mu, (size_t)mcb * m->p, dev, m->stream));
    GPB_TRY(copy_out(var ? var + c0 : nullptr, m->var, mcb, dev, m->stream));
    if (!dev) GPB_CUDA(cudaStreamSynchronize(m->stream));
  }
  GPB_CUDA(cudaStreamSynchronize(m->stream));
  return 0;
}

int gpb_model_predict_full_cov(gpb_model *m, int mc, const double *Xc, int include_likelihood, double *mu, double *cov, int dev) {
  GPB_RANGE("gpb_model_predict_full_cov");
  GPB_REQUIRE(m && Xc && cov, "predict_full_cov: NULL argument");
  GPB_TRY(check_device(m, "predict_full_cov"));
  GPB_REQUIRE(m->fitted, "predict_full_cov: model has not been fitted");
  GPB_REQUIRE(mc >= 1 && mc <= m->cb, "predict_full_cov: mc = %d exceeds the candidate block %d", mc, m->cb);
  const int cpad = round_up(mc, TILE), np = m->np;
  cudaStream_t s = m->stream;
  AllocStream alloc_scope(s);
  GPB_TRY(predict_block(m, Xc, mc, dev, 1, include_likelihood, false));
  // cov = Kxx - tmp^T tmp (+ noise I)      posterior.py:281-284, gaussian.py:104-107
  DevBuf cbuf;
  GPB_TRY(cbuf.alloc((size_t)cpad * cpad * sizeof(double)));
  const KCoords kc = train_coords(m);
  const double *XcK = m->gower ? m->XcgT : m->XcT;
  GPB_TRY(launch_kmat(m->kind, XcK, cpad, XcK, cpad, m->d, mc, mc, kc.var, include_likelihood ? m->noise : 0.0, 1, cbuf.d(), cpad,
                      cpad, cpad, s, kc.gflag));
  GemmArgs g{m->Vt, np, m->Vt, np, cbuf.d(), cpad, cpad, cpad, np, -1.0, 1.0, 0, 0, 0};
  GPB_TRY(gemm_launch(LAYOUT_ROWK, LAYOUT_ROWK, g, s));
  GPB_TRY(copy_out(mu, m->mu, (size_t)mc * m->p, dev, s));
  GPB_CUDA(cudaMemcpy2DAsync(cov, (size_t)mc * sizeof(double), cbuf.d(), (size_t)cpad * sizeof(double), (size_t)mc * sizeof(double), mc,
                             dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, s));
  GPB_CUDA(cudaStreamSynchronize(s));
  return 0;
}

int gpb_model_predictive_gradients(gpb_model *m, int mc, const double *Xc, double *dmu, double *dvar, int dev) {
  GPB_RANGE("gpb_model_predictive_gradients");
  GPB_REQUIRE(m && Xc, "predictive_gradients: NULL argument");
  GPB_TRY(check_device(m, "predictive_gradients"));
  GPB_REQUIRE(m->fitted, "predictive_gradients: model has not been fitted");
  for (int c0 = 0; c0 < mc; c0 += m->cb) {
    const int mcb = std::min(m->cb, mc - c0);
    GPB_TRY(predict_block(m, Xc + (size_t)c0 * m->d, mcb, dev, dvar ? 2 : 3, 0));
    GPB_TRY(copy_out(dmu ? dmu + (size_t)c0 * m->d : nullptr, m->dmu, (size_t)mcb * m->d, dev, m->stream));
    GPB_TRY(copy_out(dvar ? dvar + (size_t)c0 * m->d : nullptr, m->dvar, (size_t)mcb * m->d, dev, m->stream));
    if (!dev) GPB_CUDA(cudaStreamSynchronize(m->stream));
  }
  GPB_CUDA(cudaStreamSynchronize(m->stream));
  return 0;
}

int gpb_model_fmin(gpb_model *m, double *fmin) {
  GPB_REQUIRE(m && fmin, "fmin: NULL argument");
  GPB_TRY(check_device(m, "fmin"));
  GPB_REQUIRE(m->fitted, "fmin: model has not been fitted");
  GPB_REQUIRE(m->p == 1, "fmin: single output only");
  // model.predict(model.X)[0].min()   gpmodel.py:125-129 -- only the mean is used, so the variance solve is skipped
  double best = INFINITY;
  for (int c0 = 0; c0 < m->n; c0 += m->cb) {
    const int mcb = std::min(m->cb, m->n - c0);
    GPB_TRY(predict_block(m, m->X + (size_t)c0 * m->d, mcb, 1, 0, 1));
    GPB_TRY(launch_min(m->mu, mcb, m->scal + 0, m->stream));
    GPB_CUDA(cudaMemcpyAsync(m->pinned, m->scal, sizeof(double), cudaMemcpyDeviceToHost, m->stream));
    GPB_CUDA(cudaStreamSynchronize(m->stream));
    best = std::min(best, m->pinned[0]);
  }
  *fmin = best;
  return 0;
}

int gpb_model_acquisition(gpb_model *m, int acq, double par, double fmin, int mc, const double *Xc, double *f, double *df,
                          double *mean, double *sd, double *dmdx, double *dsdx, int dev) {
  GPB_RANGE("gpb_model_acquisition");
  GPB_REQUIRE(m && Xc, "acquisition: NULL argument");
  GPB_TRY(check_device(m, "acquisition"));
  GPB_REQUIRE(m->fitted, "acquisition: model has not been fitted");
  GPB_REQUIRE(acq == GPB_ACQ_EI || acq == GPB_ACQ_LCB, "acquisition: unknown type %d", acq);
  GPB_REQUIRE(m->p == 1, "acquisition: single output only");
  const bool grad = df || dmdx || dsdx;
  const int d = m->d;
  if (!dev && mc >= 1 && mc <= 8) {
    // the refinement call (optimizer.py:46-51): one upload, the fused kernel, the epilogue, staged results, ONE synchronisation
    GPB_TRY(predict_block(m, Xc, mc, 0, grad ? 2 : 1, 1));
    GPB_TRY(launch_acq_epilogue(acq, par, fmin, mc, d, m->mu, m->var, grad ? m->dmu : nullptr, grad ? m->dvar : nullptr, m->fbuf,
                                m->dfbuf, nullptr, m->sdbuf, nullptr, m->dsbuf, m->stream));
    GPB_TRY(copy_out_staged(m, f, m->fbuf, mc));
    GPB_TRY(copy_out_staged(m, mean, m->mu, mc));
    GPB_TRY(copy_out_staged(m, sd, m->sdbuf, mc));
    if (grad) {
      GPB_TRY(copy_out_staged(m, df, m->dfbuf, (size_t)mc * d));
      GPB_TRY(copy_out_staged(m, dmdx, m->dmu, (size_t)mc * d));
      GPB_TRY(copy_out_staged(m, dsdx, m->dsbuf, (size_t)mc * d));
    }
    return finish_staged(m);
  }
  for (int c0 = 0; c0 < mc; c0 += m->cb) {
    const int mcb = std::min(m->cb, mc - c0);
    GPB_TRY(predict_block(m, Xc + (size_t)c0 * d, mcb, dev, grad ? 2 : 1, 1));
    GPB_TRY(launch_acq_epilogue(acq, par, fmin, mcb, d, m->mu, m->var, grad ? m->dmu : nullptr, grad ? m->dvar : nullptr, m->fbuf,
                                m->dfbuf, nullptr, m->sdbuf, nullptr, m->dsbuf, m->stream));
    GPB_TRY(copy_out(f ? f + c0 : nullptr, m->fbuf, mcb, dev, m->stream));
    GPB_TRY(copy_out(mean ? mean + c0 : nullptr, m->mu, mcb, dev, m->stream));
    GPB_TRY(copy_out(sd ? sd + c0 : nullptr, m->sdbuf, mcb, dev, m->stream));
    if (grad) {
      GPB_TRY(copy_out(df ? df + (size_t)c0 * d : nullptr, m->dfbuf, (size_t)mcb * d, dev, m->stream));
      GPB_TRY(copy_out(dmdx ? dmdx + (size_t)c0 * d : nullptr, m->dmu, (size_t)mcb * d, dev, m->stream));
      GPB_TRY(copy_out(dsdx ? dsdx + (size_t)c0 * d : nullptr, m->dsbuf, (size_t)mcb * d, dev, m->stream));
    }
    if (!dev) GPB_CUDA(cudaStreamSynchronize(m->stream));
  }
  GPB_CUDA(cudaStreamSynchronize(m->stream));
  return 0;
}

int gpb_model_set_penalizers(gpb_model *m, int transform, int nb, const double *Xb, const double *r, const double *s) {
  GPB_REQUIRE(m, "set_penalizers: NULL model");
  GPB_TRY(check_device(m, "set_penalizers"));
  GPB_REQUIRE(transform == 0 || transform == 1, "set_penalizers: transform must be 0 (none) or 1 (softplus)");
  GPB_REQUIRE(nb >= 0 && (nb == 0 || (Xb && r && s)), "set_penalizers: bad arguments");
  m->lp_transform = transform;
  m->lp_nb = nb;
  if (nb == 0) return 0;
  const int d = m->d;
  if (nb > m->lp_cap) {
    if (m->lp_buf) GPB_CUDA(cudaFree(m->lp_buf));
    m->lp_buf = nullptr;
    m->lp_cap = std::max(64, 2 * nb);
    GPB_CUDA(cudaMalloc(&m->lp_buf, (size_t)m->lp_cap * (d + 2) * sizeof(double)));
  }
  cudaStream_t st = m->stream;
  GPB_CUDA(cudaMemcpyAsync(m->lp_buf, Xb, (size_t)nb * d * sizeof(double), cudaMemcpyHostToDevice, st));
  GPB_CUDA(cudaMemcpyAsync(m->lp_buf + (size_t)m->lp_cap * d, r, nb * sizeof(double), cudaMemcpyHostToDevice, st));
  GPB_CUDA(cudaMemcpyAsync(m->lp_buf + (size_t)m->lp_cap * (d + 1), s, nb * sizeof(double), cudaMemcpyHostToDevice, st));
  GPB_CUDA(cudaStreamSynchronize(st));
  return 0;
}

int gpb_model_acquisition_lp(gpb_model *m, int acq, double par, double fmin, int mc, const double *Xc, double *f, double *df, int dev) {
  GPB_RANGE("gpb_model_acquisition_lp");
  GPB_REQUIRE(m && Xc && f, "acquisition_lp: NULL argument");
  GPB_TRY(check_device(m, "acquisition_lp"));
  GPB_REQUIRE(m->fitted, "acquisition_lp: model has not been fitted");
  GPB_REQUIRE(acq == GPB_ACQ_EI || acq == GPB_ACQ_LCB, "acquisition_lp: unknown type %d", acq);
  GPB_REQUIRE(m->p == 1, "acquisition_lp: single output only");
  const int d = m->d;
  const bool grad = df != nullptr;
  const double *Xb = m->lp_buf, *r = m->lp_buf + (size_t)m->lp_cap * d, *s = m->lp_buf + (size_t)m->lp_cap * (d + 1);
  if (!dev && mc >= 1 && mc <= 8) {
    GPB_TRY(predict_block(m, Xc, mc, 0, grad ? 2 : 1, 1));
    GPB_TRY(launch_acq_epilogue(acq, par, fmin, mc, d, m->mu, m->var, grad ? m->dmu : nullptr, grad ? m->dvar : nullptr, m->fbuf,
                                m->dfbuf, nullptr, nullptr, nullptr, nullptr, m->stream));
    GPB_TRY(launch_lp_epilogue(mc, d, m->lp_nb, m->Xc, Xb, r, s, m->lp_transform, m->fbuf, grad ? m->dfbuf : nullptr, m->sdbuf,
                               grad ? m->dsbuf : nullptr, m->stream));
    GPB_TRY(copy_out_staged(m, f, m->sdbuf, mc));
    if (grad) GPB_TRY(copy_out_staged(m, df, m->dsbuf, (size_t)mc * d));
    return finish_staged(m);
  }
  for (int c0 = 0; c0 < mc; c0 += m->cb) {
    const int mcb = std::min(m->cb, mc - c0);
    GPB_TRY(predict_block(m, Xc + (size_t)c0 * d, mcb, dev, grad ? 2 : 1, 1));
    GPB_TRY(launch_acq_epilogue(acq, par, fmin, mcb, d, m->mu, m->var, grad ? m->dmu : nullptr, grad ? m->dvar : nullptr, m->fbuf,
                                m->dfbuf, nullptr, nullptr, nullptr, nullptr, m->stream));
    // fbuf / dfbuf hold -acq / -dacq; the LP epilogue rewrites them in place-compatible buffers (sdbuf, dsbuf are free here)
    GPB_TRY(launch_lp_epilogue(mcb, d, m->lp_nb, m->Xc, Xb, r, s, m->lp_transform, m->fbuf, grad ? m->dfbuf : nullptr, m->sdbuf,
                               grad ? m->dsbuf : nullptr, m->stream));
    GPB_TRY(copy_out(f + c0, m->sdbuf, mcb, dev, m->stream));
    if (grad) GPB_TRY(copy_out(df + (size_t)c0 * d, m->dsbuf, (size_t)mcb * d, dev, m->stream));
    if (!dev) GPB_CUDA(cudaStreamSynchronize(m->stream));
  }
  GPB_CUDA(cudaStreamSynchronize(m->stream));
  return 0;
}

// ---- two-lane block pipeline of the scoring pass --------------------------------------------------------------------------------
static gpb_model::Lane lane_of(const gpb_model *m) {
  return {m->Xc, m->XcT, m->XcgT, m->KxT, m->Vt, m->Ut, m->mu, m->var, m->dmu, m->dvar, m->fbuf, m->dfbuf, m->sdbuf, m->dsbuf, m->stream};
}
static void lane_activate(gpb_model *m, const gpb_model::Lane &l) {
  m->Xc = l.Xc; m->XcT = l.XcT; m->XcgT = l.XcgT; m->KxT = l.KxT; m->Vt = l.Vt; m->Ut = l.Ut; m->mu = l.mu; m->var = l.var;
  m->dmu = l.dmu; m->dvar = l.dvar; m->fbuf = l.fbuf; m->dfbuf = l.dfbuf; m->sdbuf = l.sdbuf; m->dsbuf = l.dsbuf; m->stream = l.stream;
}
static int lane_ensure(gpb_model *m) {
  if (m->lane_mem) return 0;
  const size_t cb = m->cb, d = m->d, np = m->np_cap, p = m->p;
  const size_t counts[14] = {cb * d, cb * d, cb * d, cb * np, cb * np, cb * np, cb * p, cb, cb * d, cb * d, cb, cb * d, cb, cb * d};
  size_t total = 0;
  for (size_t c : counts) total += align256(c * sizeof(double));
  GPB_CUDA(cudaMalloc(&m->lane_mem, total));
  char *base = reinterpret_cast<char *>(m->lane_mem);
  double **dst[14] = {&m->lane1.Xc, &m->lane1.XcT, &m->lane1.XcgT, &m->lane1.KxT, &m->lane1.Vt, &m->lane1.Ut, &m->lane1.mu, &m->lane1.var,
                      &m->lane1.dmu, &m->lane1.dvar, &m->lane1.fbuf, &m->lane1.dfbuf, &m->lane1.sdbuf, &m->lane1.dsbuf};
  size_t off = 0;
  for (int i = 0; i < 14; ++i) {
    *dst[i] = reinterpret_cast<double *>(base + off);
    off += align256(counts[i] * sizeof(double));
  }
  GPB_CUDA(cudaStreamCreateWithFlags(&m->lane_stream, cudaStreamNonBlocking));
  for (auto &e : m->lane_ev) GPB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  m->lane1.stream = m->lane_stream;
  return 0;
}

// One scoring pass with a running top-k on the device.  The k result rows [value, global index, coordinates] are left in
// out_rows_dev (device, k x (d + 2); needs device-resident candidates) when it is given.  No synchronisation at the end.
static int acq_topk_pass(gpb_model *m, int acq, double par, double fmin, int mc, const double *Xc, int dev, int k, long long index_offset,
                         double *f, double *df, double *out_rows_dev) {
  const int d = m->d;
  const bool grad = df != nullptr;
  cudaStream_t s0 = m->stream;
  GPB_TRY(launch_topk_init(m->topv, m->topi, k, s0));
  // Device-resident candidates in several blocks: the blocks alternate between two lanes (buffer sets + streams).  Within a lane the
  // stream orders the reuse of its buffers; across lanes only the running top-k is shared, so its updates are chained by an event.
  static const int two_lanes_on = env_int("GPB_ACQ_LANES", 2) >= 2;
  const bool two = two_lanes_on && dev && mc > m->cb && !m->gower;
  const gpb_model::Lane lane0 = lane_of(m);
  if (two) {
    GPB_TRY(lane_ensure(m));
    GPB_CUDA(cudaEventRecord(m->lane_ev[0], s0));                 // lane 1 starts behind everything queued so far (the fit, top-k init)
    GPB_CUDA(cudaStreamWaitEvent(m->lane_stream, m->lane_ev[0], 0));
  }
  int rc = 0, blk = 0;
  bool chain = false;
  for (int c0 = 0; c0 < mc && rc == 0; c0 += m->cb, ++blk) {
    const int mcb = std::min(m->cb, mc - c0);
    if (two) lane_activate(m, (blk & 1) ? m->lane1 : lane0);
    cudaStream_t s = m->stream;
    rc = predict_block(m, Xc + (size_t)c0 * d, mcb, dev, grad ? 2 : 1, 1);
    if (rc == 0)
      rc = launch_acq_epilogue(acq, par, fmin, mcb, d, m->mu, m->var, grad ? m->dmu : nullptr, grad ? m->dvar : nullptr, m->fbuf,
                               grad ? m->dfbuf : nullptr, nullptr, nullptr, nullptr, nullptr, s);
    if (rc == 0 && two && chain && cudaStreamWaitEvent(s, m->lane_ev[1], 0) != cudaSuccess) rc = -1;   // the previous block's top-k update
    if (rc == 0) rc = launch_topk_update(m->fbuf, mcb, index_offset + c0, m->topv, m->topi, k, s);
    if (rc == 0 && two) {
      if (cudaEventRecord(m->lane_ev[1], s) != cudaSuccess) rc = -1;
      chain = true;
    }
    if (rc == 0) rc = copy_out(f ? f + c0 : nullptr, m->fbuf, mcb, dev, s);
    if (rc == 0 && grad) rc = copy_out(df + (size_t)c0 * d, m->dfbuf, (size_t)mcb * d, dev, s);
    if (rc == 0 && !dev && cudaStreamSynchronize(s) != cudaSuccess) rc = -1;  // the host block may be reused by the caller's next copy
  }
  if (two) {
    lane_activate(m, lane0);
    // join: the model's stream continues behind lane 1 (and behind the last top-k update, wherever it ran)
    if (cudaEventRecord(m->lane_ev[2], m->lane_stream) != cudaSuccess || cudaStreamWaitEvent(s0, m->lane_ev[2], 0) != cudaSuccess) rc = rc ? rc : -1;
  }
  if (rc != 0) {
    if (rc == -1 && gpb_last_error()[0] == 0) set_error("acq_topk: CUDA error in the block pipeline");
    return rc;
  }
  if (out_rows_dev) GPB_TRY(launch_topk_pack(m->topv, m->topi, Xc, d, k, index_offset, out_rows_dev, s0));
  return 0;
}

static int acq_topk_check(gpb_model *m, int acq, int mc, int k, const char *what) {
  GPB_TRY(check_device(m, what));
  GPB_REQUIRE(m->fitted, "%s: model has not been fitted", what);
  GPB_REQUIRE(acq == GPB_ACQ_EI || acq == GPB_ACQ_LCB, "%s: unknown acquisition type %d", what, acq);
  GPB_REQUIRE(k >= 1 && k <= 64 && k <= mc, "%s: k = %d must be in [1, min(64, mc)]", what, k);
  GPB_REQUIRE(m->p == 1, "%s: single output only", what);
  return 0;
}

int gpb_model_acq_topk_full(gpb_model *m, int acq, double par, double fmin, int mc, const double *Xc, int dev, int k,
                            long long index_offset, double *vals, long long *idx, double *pts, double *f, double *df) {
  GPB_RANGE("gpb_model_acq_topk_full");
  GPB_REQUIRE(m && Xc && vals && idx, "acq_topk: NULL argument");
  GPB_TRY(acq_topk_check(m, acq, mc, k, "acq_topk"));
  const int d = m->d;
  cudaStream_t s = m->stream;
  AllocStream alloc_scope(s);
  GPB_TRY(acq_topk_pass(m, acq, par, fmin, mc, Xc, dev, k, index_offset, f, df, nullptr));
  GPB_CUDA(cudaMemcpyAsync(m->pinned, m->topv, k * sizeof(double), cudaMemcpyDeviceToHost, s));
  GPB_CUDA(cudaMemcpyAsync(m->pinned + 64, m->topi, k * sizeof(long long), cudaMemcpyDeviceToHost, s));
  GPB_CUDA(cudaStreamSynchronize(s));
  for (int i = 0; i < k; ++i) {
    vals[i] = m->pinned[i];
    idx[i] = reinterpret_cast<long long *>(m->pinned + 64)[i];
  }
  // A slot that never received a candidate (NaN scores never win: fewer than k finite scores) still holds the initial sentinel:
  // report it as (NaN, -1, NaN ...), the empty-slot convention sharded.merge_topk filters on, instead of dereferencing it.
  for (int i = 0; i < k; ++i) {
    const bool empty = idx[i] == LLONG_MAX;
    if (empty) {
      vals[i] = NAN;
      idx[i] = -1;
    }
    if (!pts) continue;
    if (empty) {
      for (int q = 0; q < d; ++q) pts[(size_t)i * d + q] = NAN;
    } else {
      GPB_CUDA(cudaMemcpy(pts + (size_t)i * d, Xc + (size_t)(idx[i] - index_offset) * d, d * sizeof(double),
                          dev ? cudaMemcpyDeviceToHost : cudaMemcpyHostToHost));
    }
  }
  return 0;
}

int gpb_model_acq_topk_dev(gpb_model *m, int acq, double par, double fmin, int mc, const double *Xc_dev, int k, long long index_offset,
                           double *rows_dev, double *f_dev, double *df_dev) {
  GPB_RANGE("gpb_model_acq_topk_dev");
  GPB_REQUIRE(m && Xc_dev && rows_dev, "acq_topk_dev: NULL argument");
  GPB_TRY(acq_topk_check(m, acq, mc, k, "acq_topk_dev"));
  AllocStream alloc_scope(m->stream);
  GPB_TRY(acq_topk_pass(m, acq, par, fmin, mc, Xc_dev, 1, k, index_offset, f_dev, df_dev, rows_dev));
  return leave_stream_ordered(m);
}

// ---- multi-GPU state distribution -------------------------------------------------------------------------------------------
int gpb_model_state_ptr(gpb_model *m, const char *what, void **ptr, size_t *count) {
  GPB_REQUIRE(m && what && ptr && count, "state_ptr: NULL argument");
  GPB_TRY(check_device(m, "state_ptr"));
  const std::string w(what);
  const size_t np = m->np;
  if (w == "Li") {
    *ptr = m->f.Mi;
    *count = np * np;
  } else if (w == "L") {
    if (m->fitted) {
      GPB_TRY(factor_finalize_L(m->f));
      GPB_CUDA(cudaStreamSynchronize(m->stream));
    }
    *ptr = m->f.A;
    *count = np * np;
  } else if (w == "Wi") {
    if (m->fitted) {
      GPB_TRY(factor_finalize_L(m->f));     // W doubles as the parking place of L's off-diagonal blocks
      GPB_TRY(ensure_wi(m));
      GPB_CUDA(cudaStreamSynchronize(m->stream));
    }
    *ptr = m->f.W;
    *count = np * np;
  } else if (w == "alpha") {
    *ptr = m->alpha;
    *count = np * (size_t)m->p;
  } else {
    GPB_REQUIRE(false, "state_ptr: unknown region '%s' (Li, L, Wi, alpha)", what);
  }
  return 0;
}

int gpb_model_adopt_state(gpb_model *m, double variance, const double *lengthscale, double noise, double jitter, int have_mask) {
  GPB_RANGE("gpb_model_adopt_state");
  GPB_REQUIRE(m && lengthscale, "adopt_state: NULL argument");
  GPB_TRY(check_device(m, "adopt_state"));
  GPB_REQUIRE(m->have_data, "adopt_state: set_data has not been called (every rank holds the same X, Y)");
  m->variance = variance;
  m->noise = noise;
  m->jitter = jitter;
  for (int q = 0; q < m->nls; ++q) m->ls[q] = lengthscale[q];
  m->scaled_valid = false;
  GPB_TRY(ensure_scaled(m));                 // scaled inputs and the lengthscale arrays the predictive kernels read
  ozaki_invalidate();                        // cached digit planes of the previous L^-1 are stale
  m->f.l_pending = false;
  m->f.l_from = 0;
  m->wi_from = 0;
  m->have_wi = (have_mask & 2) != 0;
  m->l_valid = (have_mask & 1) != 0;
  m->fitted = true;
  GPB_CUDA(cudaStreamSynchronize(m->stream));
  return 0;
}

int gpb_model_acq_topk(gpb_model *m, int acq, double par, double fmin, int mc, const double *Xc, int dev, int k,
                       long long index_offset, double *vals, long long *idx, double *pts) {
  GPB_RANGE("gpb_model_acq_topk");
  return gpb_model_acq_topk_full(m, acq, par, fmin, mc, Xc, dev, k, index_offset, vals, idx, pts, nullptr, nullptr);
}

// =====================================================================================================================
// Kern contract (stateless)
// =====================================================================================================================
struct KernTmp {
  DevBuf xa, xb, ls, xat, xbt;
  const double *Xa = nullptr, *Xb = nullptr;
  double *XaT = nullptr, *XbT = nullptr, *ls_dev = nullptr, *inv_ls_dev = nullptr;
  int npa = 0, npb = 0;
};

static int kern_prepare(KernTmp &t, int d, int n, const double *X, int m, const double *X2, const double *ls, int nls, int dev,
                        cudaStream_t s) {
  GPB_REQUIRE(X && ls && d >= 1 && d <= 96 && n >= 1, "kern: bad arguments");
  GPB_REQUIRE(nls == 1 || nls == d, "kern: lengthscale must have 1 or input_dim entries (got %d)", nls);
  GPB_REQUIRE(gpb_device_count() > 0, "kern: no CUDA device visible -- libgpb200 has no CPU fallback");
  t.npa = round_up(n, TILE);
  GPB_TRY(to_device(t.xa, X, (size_t)n * d, dev, &t.Xa, s));
  GPB_TRY(t.ls.alloc(2 * d * sizeof(double)));
  t.ls_dev = t.ls.d();
  t.inv_ls_dev = t.ls.d() + d;
  GPB_TRY(set_ls(ls, nls, d, t.ls_dev, t.inv_ls_dev, s));
  GPB_TRY(t.xat.alloc((size_t)d * t.npa * sizeof(double)));
  t.XaT = t.xat.d();
  GPB_TRY(launch_scale_transpose(t.Xa, n, d, t.ls_dev, t.XaT, t.npa, s));
  if (X2) {
    GPB_REQUIRE(m >= 1, "kern: X2 given with m = %d", m);
    t.npb = round_up(m, TILE);
    GPB_TRY(to_device(t.xb, X2, (size_t)m * d, dev, &t.Xb, s));
    GPB_TRY(t.xbt.alloc((size_t)d * t.npb * sizeof(double)));
    t.XbT = t.xbt.d();
    GPB_TRY(launch_scale_transpose(t.Xb, m, d, t.ls_dev, t.XbT, t.npb, s));
  } else {
    t.npb = t.npa;
    t.XbT = t.XaT;
  }
  return 0;
}

int gpb_kern_K(int kind, int d, int n, const double *X, int m, const double *X2, double variance, const double *lengthscale,
               int nls, double *K, int ldk, int dev, void *stream) {
  GPB_RANGE("gpb_kern_K");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  AllocStream alloc_scope(s);
  GPB_REQUIRE(K != nullptr, "kern_K: K is NULL");
  GPB_REQUIRE(kind == GPB_KERN_RBF || kind == GPB_KERN_MATERN52, "kern_K: unknown kernel kind %d", kind);
  KernTmp t;
  GPB_TRY(kern_prepare(t, d, n, X, m, X2, lengthscale, nls, dev, s));
  const int cols = X2 ? m : n;
  GPB_REQUIRE(ldk >= cols, "kern_K: ldk too small");
  DevBuf out;
  double *Kd = K;
  int ld = ldk;
  if (!dev) {
    GPB_TRY(out.alloc((size_t)n * cols * sizeof(double)));
    Kd = out.d();
    ld = cols;
  }
  GPB_TRY(launch_kmat(kind, t.XaT, t.npa, t.XbT, t.npb, d, n, cols, variance, 0.0, 0, Kd, ld, t.npa, t.npb, s));
  if (!dev)
    GPB_CUDA(cudaMemcpy2DAsync(K, (size_t)ldk * sizeof(double), Kd, (size_t)ld * sizeof(double), (size_t)cols * sizeof(double), n,
                               cudaMemcpyDeviceToHost, s));
  GPB_CUDA(cudaStreamSynchronize(s));
  return 0;
}

// Gower coordinates for the stateless Kern calls: XaT / XbT = x / range (continuous) or x (discrete), flags on the device
struct GowerTmp {
  DevBuf xa, xb, xat, xbt, par;
  double *XaT = nullptr, *XbT = nullptr, *gflag = nullptr;
  int npa = 0, npb = 0;
};

static int gower_prepare(GowerTmp &t, int d, int n, const double *X, int m, const double *X2, const int *discrete, const double *ranges,
                         int dev, cudaStream_t s) {
  GPB_REQUIRE(X && discrete && ranges && d >= 1 && d <= 96 && n >= 1, "kern (Gower): bad arguments");
  GPB_REQUIRE(gpb_device_count() > 0, "kern: no CUDA device visible -- libgpb200 has no CPU fallback");
  std::vector<double> h(2 * d);
  for (int q = 0; q < d; ++q) {
    h[q] = discrete[q] ? 1.0 : 0.0;
    h[d + q] = discrete[q] ? 1.0 : ranges[q];
    GPB_REQUIRE(h[d + q] > 0, "kern (Gower): range of continuous dimension %d must be positive", q);
  }
  GPB_TRY(t.par.alloc(2 * d * sizeof(double)));
  GPB_CUDA(cudaMemcpyAsync(t.par.p, h.data(), 2 * d * sizeof(double), cudaMemcpyHostToDevice, s));
  GPB_CUDA(cudaStreamSynchronize(s));
  t.gflag = t.par.d();
  const double *scale = t.par.d() + d, *Xa = nullptr, *Xb = nullptr;
  t.npa = round_up(n, TILE);
  GPB_TRY(to_device(t.xa, X, (size_t)n * d, dev, &Xa, s));
  GPB_TRY(t.xat.alloc((size_t)d * t.npa * sizeof(double)));
  t.XaT = t.xat.d();
  GPB_TRY(launch_scale_transpose(Xa, n, d, scale, t.XaT, t.npa, s, 1));
  if (X2) {
    GPB_REQUIRE(m >= 1, "kern: X2 given with m = %d", m);
    t.npb = round_up(m, TILE);
    GPB_TRY(to_device(t.xb, X2, (size_t)m * d, dev, &Xb, s));
    GPB_TRY(t.xbt.alloc((size_t)d * t.npb * sizeof(double)));
    t.XbT = t.xbt.d();
    GPB_TRY(launch_scale_transpose(Xb, m, d, scale, t.XbT, t.npb, s, 1));
  } else {
    t.npb = t.npa;
    t.XbT = t.XaT;
  }
  return 0;
}

int gpb_kern_K_gower(int kind, int d, int n, const double *X, int m, const double *X2, double variance, const int *discrete,
                     const double *ranges, double *K, int ldk, int dev, void *stream) {
  GPB_RANGE("gpb_kern_K_gower");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  AllocStream alloc_scope(s);
  GPB_REQUIRE(K != nullptr, "kern_K: K is NULL");
  GPB_REQUIRE(kind == GPB_KERN_RBF || kind == GPB_KERN_MATERN52, "kern_K: unknown kernel kind %d", kind);
  GowerTmp t;
  GPB_TRY(gower_prepare(t, d, n, X, m, X2, discrete, ranges, dev, s));
  const int cols = X2 ? m : n;
  GPB_REQUIRE(ldk >= cols, "kern_K: ldk too small");
  DevBuf out;
  double *Kd = K;
  int ld = ldk;
  if (!dev) {
    GPB_TRY(out.alloc((size_t)n * cols * sizeof(double)));
    Kd = out.d();
    ld = cols;
  }
  GPB_TRY(launch_kmat(kind, t.XaT, t.npa, t.XbT, t.npb, d, n, cols, std::pow(variance, d), 0.0, 0, Kd, ld, t.npa, t.npb, s, t.gflag));
  if (!dev)
    GPB_CUDA(cudaMemcpy2DAsync(K, (size_t)ldk * sizeof(double), Kd, (size_t)ld * sizeof(double), (size_t)cols * sizeof(double), n,
                               cudaMemcpyDeviceToHost, s));
  GPB_CUDA(cudaStreamSynchronize(s));
  return 0;
}

int gpb_kern_update_gradients_full(int kind, int d, int n, const double *X, int m, const double *X2, const double *dL_dK, int ld,
                                   double variance, const double *lengthscale, int nls, double *out, int dev, void *stream);

int gpb_kern_update_gradients_full_gower(int kind, int d, int n, const double *X, int m, const double *X2, const double *dL_dK, int ld,
                                         double variance, const double *lengthscale, int nls, const int *discrete,
                                         const double *ranges, double *out, int dev, void *stream) {
  GPB_RANGE("gpb_kern_update_gradients_full_gower");
  // lengthscale terms: the reference keeps the Euclidean distance under the patch (stationary.py:227-238)
  GPB_TRY(gpb_kern_update_gradients_full(kind, d, n, X, m, X2, dL_dK, ld, variance, lengthscale, nls, out, dev, stream));
  // variance term: sum(K_gower * dL_dK) / variance (stationary.py:224 with the patched K)
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  AllocStream alloc_scope(s);
  GowerTmp t;
  GPB_TRY(gower_prepare(t, d, n, X, m, X2, discrete, ranges, dev, s));
  const int cols = X2 ? m : n;
  DevBuf gbuf, part, res;
  const double *G = dL_dK;
  int ldg = ld;
  if (!dev) {
    GPB_TRY(gbuf.alloc((size_t)n * cols * sizeof(double)));
    GPB_CUDA(cudaMemcpy2DAsync(gbuf.p, (size_t)cols * sizeof(double), dL_dK, (size_t)ld * sizeof(double), (size_t)cols * sizeof(double), n,
                               cudaMemcpyHostToDevice, s));
    G = gbuf.d();
    ldg = cols;
  }
  GPB_TRY(part.alloc(kgrad_part_doubles(t.npa, t.npb, d, 0) * sizeof(double)));
  GPB_TRY(res.alloc(sizeof(double)));
  GPB_TRY(launch_kvar_gower(kind, 0, t.XaT, t.npa, t.XbT, t.npb, d, n, cols, std::pow(variance, d), t.gflag, G, ldg, nullptr, 0, 1,
                            part.d(), res.d(), s));
  double h = 0.0;
  GPB_CUDA(cudaMemcpyAsync(&h, res.p, sizeof(double), cudaMemcpyDeviceToHost, s));
  GPB_CUDA(cudaStreamSynchronize(s));
  out[0] = h / variance;
  return 0;
}

int gpb_kern_update_gradients_full(int kind, int d, int n, const double *X, int m, const double *X2, const double *dL_dK, int ld,
                                   double variance, const double *lengthscale, int nls, double *out, int dev, void *stream) {
  GPB_RANGE("gpb_kern_update_gradients_full");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  AllocStream alloc_scope(s);
  GPB_REQUIRE(dL_dK && out, "update_gradients_full: NULL argument");
  GPB_REQUIRE(kind == GPB_KERN_RBF || kind == GPB_KERN_MATERN52, "update_gradients_full: unknown kernel kind %d", kind);
  KernTmp t;
  GPB_TRY(kern_prepare(t, d, n, X, m, X2, lengthscale, nls, dev, s));
  const int cols = X2 ? m : n;
  GPB_REQUIRE(ld >= cols, "update_gradients_full: ld too small");
  DevBuf gbuf, part, res;
  const double *G = dL_dK;
  int ldg = ld;
  if (!dev) {
    GPB_TRY(gbuf.alloc((size_t)n * cols * sizeof(double)));
    GPB_CUDA(cudaMemcpy2DAsync(gbuf.p, (size_t)cols * sizeof(double), dL_dK, (size_t)ld * sizeof(double), (size_t)cols * sizeof(double), n,
                               cudaMemcpyHostToDevice, s));
    G = gbuf.d();
    ldg = cols;
  }
  GPB_TRY(part.alloc(kgrad_part_doubles(t.npa, t.npb, d, 0) * sizeof(double)));
  GPB_TRY(res.alloc((d + 2) * sizeof(double)));
  GPB_TRY(launch_kgrad(kind, 0, t.XaT, t.npa, t.XbT, t.npb, d, n, cols, variance, G, ldg, nullptr, 0, 1, part.d(), res.d(), s));
  std::vector<double> h(d + 2);
  GPB_CUDA(cudaMemcpyAsync(h.data(), res.p, (d + 2) * sizeof(double), cudaMemcpyDeviceToHost, s));
  GPB_CUDA(cudaStreamSynchronize(s));
  out[0] = h[0] / variance;
  if (nls == d) {
    for (int q = 0; q < d; ++q) out[1 + q] = -h[2 + q] / lengthscale[q];  // stationary.py:233,268 (scaled coordinates)
  } else {
    double acc = 0.0;
    for (int q = 0; q < d; ++q) acc += h[2 + q];
    out[1] = -acc / lengthscale[0];                                       // stationary.py:237-238
  }
  return 0;
}

int gpb_kern_gradients_X(int kind, int d, int n, const double *X, int m, const double *X2, const double *dL_dK, int ld,
                         double variance, const double *lengthscale, int nls, double *out, int dev, void *stream) {
  GPB_RANGE("gpb_kern_gradients_X");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  AllocStream alloc_scope(s);
  GPB_REQUIRE(dL_dK && out, "gradients_X: NULL argument");
  GPB_REQUIRE(kind == GPB_KERN_RBF || kind == GPB_KERN_MATERN52, "gradients_X: unknown kernel kind %d", kind);
  KernTmp t;
  GPB_TRY(kern_prepare(t, d, n, X, m, X2, lengthscale, nls, dev, s));
  const int cols = X2 ? m : n;
  GPB_REQUIRE(ld >= cols, "gradients_X: ld too small");
  DevBuf gbuf, obuf;
  const double *G = dL_dK;
  int ldg = ld;
  if (!dev) {
    GPB_TRY(gbuf.alloc((size_t)n * cols * sizeof(double)));
    GPB_CUDA(cudaMemcpy2DAsync(gbuf.p, (size_t)cols * sizeof(double), dL_dK, (size_t)ld * sizeof(double), (size_t)cols * sizeof(double), n,
                               cudaMemcpyHostToDevice, s));
    G = gbuf.d();
    ldg = cols;
  }
  double *od = out;
  if (!dev) {
    GPB_TRY(obuf.alloc((size_t)n * d * sizeof(double)));
    od = obuf.d();
  }
  GPB_TRY(launch_gradx(kind, t.XaT, t.npa, n, t.XbT, t.npb, cols, d, variance, t.inv_ls_dev, G, ldg, 1.0, X2 ? 0 : 1, nullptr, 0, 0.0,
                       od, nullptr, d, s));
  if (!dev) GPB_CUDA(cudaMemcpyAsync(out, od, (size_t)n * d * sizeof(double), cudaMemcpyDeviceToHost, s));
  GPB_CUDA(cudaStreamSynchronize(s));
  return 0;
}

// =====================================================================================================================
// util.linalg
// =====================================================================================================================
int gpb_pdinv(int n, const double *A, int lda, double *L, double *Ai, double *Li, double *logdet, int dev, void *stream) {
  GPB_RANGE("gpb_pdinv");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  AllocStream alloc_scope(s);
  GPB_REQUIRE(A && n >= 1 && lda >= n, "pdinv: bad arguments");
  GPB_REQUIRE(gpb_device_count() > 0, "pdinv: no CUDA device visible -- libgpb200 has no CPU fallback");
  const int np = round_up(n, TILE), nb = np / TILE;
  DevBuf ain, a, mi, w, part, misc, outb;
  const double *Ad = A;
  int ldad = lda;
  if (!dev) {
    GPB_TRY(ain.alloc((size_t)n * n * sizeof(double)));
    GPB_CUDA(cudaMemcpy2DAsync(ain.p, (size_t)n * sizeof(double), A, (size_t)lda * sizeof(double), (size_t)n * sizeof(double), n,
                               cudaMemcpyHostToDevice, s));
    Ad = ain.d();
    ldad = n;
  }
  GPB_TRY(a.alloc((size_t)np * np * sizeof(double)));
  GPB_TRY(mi.alloc((size_t)np * np * sizeof(double)));
  GPB_TRY(w.alloc((size_t)np * np * sizeof(double)));
  GPB_TRY(part.alloc((size_t)nb * np * sizeof(double)));
  GPB_TRY(misc.alloc(64));
  Factor f;
  f.n = n;
  f.np = np;
  f.A = a.d();
  f.Mi = mi.d();
  f.W = w.d();
  f.part = part.d();
  f.info = reinterpret_cast<int *>(misc.d() + 1);
  f.stream = s;
  pad_sym_kernel<<<dim3((np + 255) / 256, np), 256, 0, s>>>(Ad, ldad, n, f.A, np);
  count_launch();
  GPB_CHECK_LAUNCH();
  GPB_TRY(factor_potrf_inv(f));
  GPB_TRY(factor_logdet(f, misc.d()));
  double h[2];
  GPB_CUDA(cudaMemcpyAsync(h, misc.p, 2 * sizeof(double), cudaMemcpyDeviceToHost, s));
  GPB_CUDA(cudaStreamSynchronize(s));
  const int info = *reinterpret_cast<int *>(&h[1]);
  if (info != 0) {
    set_error("pdinv: matrix not positive definite (leading minor %d)", info);
    return info > n ? n : info;
  }
  if (logdet) *logdet = h[0];
  GPB_TRY(factor_finalize_L(f));
  if (Ai) GPB_TRY(factor_potri(f));
  if (!dev && (L || Ai || Li)) GPB_TRY(outb.alloc((size_t)n * n * sizeof(double)));
  const dim3 grid((n + 255) / 256, n), gsym((n + 31) / 32, (n + 31) / 32);
  for (int which = 0; which < 3; ++which) {
    double *dst = which == 0 ? L : which == 1 ? Li : Ai;
    if (!dst) continue;
    double *dd = dev ? dst : outb.d();
    if (which == 0)
      tril_copy_kernel<<<grid, 256, 0, s>>>(f.A, np, dd, n, n);
    else if (which == 1)
      tril_copy_kernel<<<grid, 256, 0, s>>>(f.Mi, np, dd, n, n);
    else
      sym_copy_kernel<<<gsym, dim3(32, 8), 0, s>>>(f.W, np, dd, n, n);
    count_launch();
    GPB_CHECK_LAUNCH();
    if (!dev) GPB_CUDA(cudaMemcpyAsync(dst, dd, (size_t)n * n * sizeof(double), cudaMemcpyDeviceToHost, s));
    GPB_CUDA(cudaStreamSynchronize(s));
  }
  return 0;
}

int gpb_potrs(int n, const double *L, int ldl, double *B, int nrhs, int dev, void *stream) {
  GPB_RANGE("gpb_potrs");
  // (L L^T) X = B as X = M^T (M B) with M = L^-1 rebuilt by the triangular-inverse recursion (factor_trtri).
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  AllocStream alloc_scope(s);
  GPB_REQUIRE(L && B && n >= 1 && ldl >= n && nrhs >= 1, "potrs: bad arguments");
  GPB_REQUIRE(gpb_device_count() > 0, "potrs: no CUDA device visible -- libgpb200 has no CPU fallback");
  const int np = round_up(n, TILE), nb = np / TILE;
  DevBuf lin, a, mi, w, part, misc, bbuf, yc, zc, xc;
  const double *Ld = L;
  int ldd = ldl;
  if (!dev) {
    GPB_TRY(lin.alloc((size_t)n * n * sizeof(double)));
    GPB_CUDA(cudaMemcpy2DAsync(lin.p, (size_t)n * sizeof(double), L, (size_t)ldl * sizeof(double), (size_t)n * sizeof(double), n,
                               cudaMemcpyHostToDevice, s));
    Ld = lin.d();
    ldd = n;
  }
  GPB_TRY(a.alloc((size_t)np * np * sizeof(double)));
  GPB_TRY(mi.alloc((size_t)np * np * sizeof(double)));
  GPB_TRY(w.alloc((size_t)np * np * sizeof(double)));
  GPB_TRY(part.alloc((size_t)nb * np * sizeof(double)));
  GPB_TRY(misc.alloc(64));
  // A-buffer = [tril(L) 0; 0 I]  (pad_sym mirrors the lower triangle; only the lower part is read afterwards)
  pad_sym_kernel<<<dim3((np + 255) / 256, np), 256, 0, s>>>(Ld, ldd, n, a.d(), np);
  count_launch();
  GPB_CHECK_LAUNCH();
  Factor f;
  f.n = n;
  f.np = np;
  f.A = a.d();
  f.Mi = mi.d();
  f.W = w.d();
  f.part = part.d();
  f.info = reinterpret_cast<int *>(misc.d() + 1);
  f.stream = s;
  GPB_TRY(factor_trtri(f));
  const double *Bd = nullptr;
  GPB_TRY(to_device(bbuf, B, (size_t)n * nrhs, dev, &Bd, s));
  GPB_TRY(yc.alloc((size_t)np * nrhs * sizeof(double)));
  GPB_TRY(zc.alloc((size_t)np * nrhs * sizeof(double)));
  GPB_TRY(xc.alloc((size_t)np * nrhs * sizeof(double)));
  pack_cols_kernel<<<(nrhs * np + 255) / 256, 256, 0, s>>>(Bd, n, nrhs, np, yc.d());
  count_launch();
  GPB_CHECK_LAUNCH();
  GPB_TRY(factor_solve(f, yc.d(), nrhs, zc.d(), xc.d()));
  double *outd = dev ? B : const_cast<double *>(Bd);
  unpack_cols_kernel<<<(n * nrhs + 255) / 256, 256, 0, s>>>(xc.d(), n, nrhs, np, outd);
  count_launch();
  GPB_CHECK_LAUNCH();
  if (!dev) GPB_CUDA(cudaMemcpyAsync(B, outd, (size_t)n * nrhs * sizeof(double), cudaMemcpyDeviceToHost, s));
  int info = 0;
  GPB_CUDA(cudaMemcpyAsync(&info, f.info, sizeof(int), cudaMemcpyDeviceToHost, s));
  GPB_CUDA(cudaStreamSynchronize(s));
  if (info != 0) {
    set_error("potrs: factor is singular (pivot %d)", info);
    return info;
  }
  return 0;
}

int gpb_potri(int n, const double *L, int ldl, double *Ai, int ldai, int dev, void *stream) {
  GPB_RANGE("gpb_potri");
  // A^-1 = L^-T L^-1 from the Cholesky factor: triangular-inverse recursion (dtrtri) + the lower-tile product M^T M
  // (dlauum), then mirrored like GPy's symmetrify.
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  AllocStream alloc_scope(s);
  GPB_REQUIRE(L && Ai && n >= 1 && ldl >= n && ldai >= n, "potri: bad arguments");
  GPB_REQUIRE(gpb_device_count() > 0, "potri: no CUDA device visible -- libgpb200 has no CPU fallback");
  const int np = round_up(n, TILE);
  DevBuf lin, a, mi, w, misc, outb;
  const double *Ld = L;
  int ldd = ldl;
  if (!dev) {
    GPB_TRY(lin.alloc((size_t)n * n * sizeof(double)));
    GPB_CUDA(cudaMemcpy2DAsync(lin.p, (size_t)n * sizeof(double), L, (size_t)ldl * sizeof(double), (size_t)n * sizeof(double), n,
                               cudaMemcpyHostToDevice, s));
    Ld = lin.d();
    ldd = n;
  }
  GPB_TRY(a.alloc((size_t)np * np * sizeof(double)));
  GPB_TRY(mi.alloc((size_t)np * np * sizeof(double)));
  GPB_TRY(w.alloc((size_t)np * np * sizeof(double)));
  GPB_TRY(misc.alloc(64));
  pad_sym_kernel<<<dim3((np + 255) / 256, np), 256, 0, s>>>(Ld, ldd, n, a.d(), np);
  count_launch();
  GPB_CHECK_LAUNCH();
  Factor f;
  f.n = n;
  f.np = np;
  f.A = a.d();
  f.Mi = mi.d();
  f.W = w.d();
  f.info = reinterpret_cast<int *>(misc.d() + 1);
  f.stream = s;
  GPB_TRY(factor_trtri(f));
  GPB_TRY(factor_potri(f));
  double *dd = Ai;
  int ldo = ldai;
  if (!dev) {
    GPB_TRY(outb.alloc((size_t)n * n * sizeof(double)));
    dd = outb.d();
    ldo = n;
  }
  sym_copy_kernel<<<dim3((n + 31) / 32, (n + 31) / 32), dim3(32, 8), 0, s>>>(f.W, np, dd, ldo, n);
  count_launch();
  GPB_CHECK_LAUNCH();
  if (!dev)
    GPB_CUDA(cudaMemcpy2DAsync(Ai, (size_t)ldai * sizeof(double), dd, (size_t)ldo * sizeof(double), (size_t)n * sizeof(double), n,
                               cudaMemcpyDeviceToHost, s));
  int info = 0;
  GPB_CUDA(cudaMemcpyAsync(&info, f.info, sizeof(int), cudaMemcpyDeviceToHost, s));
  GPB_CUDA(cudaStreamSynchronize(s));
  if (info != 0) {
    set_error("potri: factor is singular (pivot %d)", info);
    return info;
  }
  return 0;
}

int gpb_set_overlap(int min_n) {
  g_overlap_min_n = min_n < 0 ? 0 : min_n;
  ++g_config_epoch;
  return 0;
}
int gpb_profile_gemm(int enable) { return gemm_profile_enable(enable); }
int gpb_gemm_config(int cfg) {
  ++g_config_epoch;
  return gemm_force_config(cfg);
}
int gpb_profile_gemm_collect(double *ms, double *flops, long long *launches) { return gemm_profile_collect(ms, flops, launches); }
int gpb_profile_gemm_last(double *ms, double *flops) { return gemm_profile_last(ms, flops); }

int gpb_dgemm(int ta, int tb, int m, int n, int k, double alpha, const double *A, int lda, const double *B, int ldb, double beta,
              double *C, int ldc, void *stream) {
  GPB_RANGE("gpb_dgemm");
  GPB_REQUIRE(A && B && C, "dgemm: NULL argument");
  GemmArgs g{A, lda, B, ldb, C, ldc, m, n, k, alpha, beta, 0, 0, 0};
  return gemm_launch(ta ? LAYOUT_COLK : LAYOUT_ROWK, tb ? LAYOUT_COLK : LAYOUT_ROWK, g, reinterpret_cast<cudaStream_t>(stream));
}

// fp64 product through the int8 tensor cores (gpb_ozaki.cu); experimental, see include/gpb200.h
int gpb_ozaki_dgemm(int ta, int tb, int m, int n, int k, double alpha, const double *A, int lda, const double *B, int ldb,
                    double beta, double *C, int ldc, int tri_out, int klo_mode, int khi_mode, int tri_a, int tri_b, int slices,
                    void *stream) {
  GPB_RANGE("gpb_ozaki_dgemm");
  GPB_REQUIRE(A && B && C, "ozaki_dgemm: NULL argument");
  GemmArgs g{A, lda, B, ldb, C, ldc, m, n, k, alpha, beta, tri_out, klo_mode, khi_mode};
  return ozaki_gemm_launch(ta ? LAYOUT_COLK : LAYOUT_ROWK, tb ? LAYOUT_COLK : LAYOUT_ROWK, g, tri_a, tri_b, slices,
                           reinterpret_cast<cudaStream_t>(stream));
}
int gpb_set_ozaki(int min_n, int slices) { return ozaki_configure(min_n, slices); }
int gpb_model_engine_report(gpb_model *m, int *engine_used, double *residual) {
  GPB_REQUIRE(m, "engine_report: NULL model");
  if (engine_used) *engine_used = m->engine_used ? 1 : 0;
  if (residual) *residual = m->last_residual;
  return 0;
}
long long gpb_ozaki_fallback_count(void) { return g_engine_fallbacks.load(std::memory_order_relaxed); }
// modular mode of the int8 engine: operand bits, and the host restatement of its integer arithmetic (test hooks, no GPU needed)
int gpb_ozaki_crt_bits(int nmod, long long k) { return ozaki_crt_bits(nmod, k); }
int gpb_ozaki_crt_host_residues(const double *A, int rows, int k, int nmod, int beta, signed char *planes, double *scale) {
  return ozaki_crt_host_residues(A, rows, k, nmod, beta, planes, scale);
}
int gpb_ozaki_crt_host_combine(const int *sums, long long count, int nmod, double *X) {
  return ozaki_crt_host_combine(sums, count, nmod, X);
}

}  // extern "C"
