// c2b_api.cu — the extern "C" boundary of libcity2ba_cuda.so (include/city2ba_cuda.h) and the
// host-side orchestration of the kernels: upload -> (grid build) -> cull -> radix sort ->
// warp-cooperative BVH traversal -> compaction -> download.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#if defined(__linux__)
#include <sys/mman.h>
#endif

#include "c2b_bvh.cuh"
#include "c2b_common.cuh"
#include "c2b_compact.cuh"
#include "c2b_cull.cuh"
#include "c2b_fused.cuh"
#include "c2b_internal.h"
#include "c2b_math.cuh"
#include "c2b_noise.cuh"
#include "c2b_sample.cuh"
#include "c2b_sort.cuh"
#include "c2b_traverse.cuh"

using namespace c2b;

namespace {

// internal status of the resident pipeline: the camera batch holds more than 2^32 pairs / candidates.
// c2b_visibility_graph halves the batch and retries; the public resident entry reports C2B_ERR_INVALID.
constexpr int ERR_TOO_LARGE = -100;

inline int bit_length(uint64_t v) {
  int b = 0;
  while (v) {
    ++b;
    v >>= 1;
  }
  return b;
}

// smallest double T with sqrt_rn(T) >= max_dist, so that  m2 < T  <=>  sqrt_rn(m2) < max_dist
// (sqrt_rn is monotone non-decreasing).  NaN / non-positive max_dist admit nothing.
double exact_sq_threshold(double max_dist) {
  if (!(max_dist > 0.0)) return 0.0;  // m2 < 0 is never true (and NaN compares false)
  if (std::isinf(max_dist)) return INFINITY;
  double x = max_dist * max_dist;
  if (std::isinf(x)) x = 1.7976931348623157e308;
  while (x > 0.0 && std::sqrt(x) >= max_dist) x = std::nextafter(x, 0.0);
  while (std::sqrt(x) < max_dist) {
    double nx = std::nextafter(x, INFINITY);
    if (std::isinf(nx)) return nx;
    x = nx;
  }
  return x;
}

// Small read-backs (counters) go through a kernel that stores into mapped pinned host memory, not
// through cudaMemcpy: a memcpy would queue behind the bulk result transfer on the D2H copy engine
// and stall the next camera batch.
__global__ void k_copy_words(const uint32_t *__restrict__ src, uint32_t *__restrict__ dst, int n) {
  if ((int)threadIdx.x < n) dst[threadIdx.x] = src[threadIdx.x];
}

__global__ void k_iota_u32(uint32_t *v, uint64_t n) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = (uint32_t)i;
}

inline unsigned blocks_for(uint64_t n, unsigned threads) { return (unsigned)((n + threads - 1) / threads); }

struct GridCache {
  bool valid = false;
  double max_dist = 0;
  uint64_t points_version = 0;
  GridDesc desc;
  uint64_t n_cells = 0;
  // hinted: the cells cover only what the cameras of that call could reach (their centres' box grown by
  // max_dist), because cells of side max_dist/4 over the bounding box of ALL points would have exceeded the
  // cell budget (outliers); such a grid serves a later call only if its cameras stay inside the hint
  bool hinted = false;
  double hint_lo[3] = {0, 0, 0}, hint_hi[3] = {0, 0, 0};
};

// Measurement / test hooks (include/city2ba_cuda.h, c2b_tune).  Defaults are the product's behaviour; the
// environment variables of the same names in upper case with a C2B_ prefix are read ONCE, in c2b_init.
struct Tunables {
  int parts_log2 = -1;             // tickets per camera = 2^k; -1 = automatic
  int trilist_cap = 128;           // entries of a camera's leaf list before it counts as overflowed
  uint64_t max_pairs = 0xffffffffull;  // 32-bit scratch offsets: pairs inside max_dist per resident pass
  int hoist_max = 64;              // leaf lists up to this length get per-camera triangle records (<= FU_HOIST)
  int packet_bvh = 1;              // list overflow: 1 = lane = node packet traversal, 0 = per-ray stackless walk
  int trilist_warp = -1;           // leaf lists by one warp per camera (1) / one thread (0); -1 = by camera count
  int batches = 0;                 // camera batches of the host-buffer call; 0 = automatic
  int fu_occ3 = 0;                 // fused kernel at 3 CTAs/SM (80 registers) instead of 4 (64)
  double grid_cell_factor = 0.25;  // point-grid cell side / max_dist
  int stage_threads = 4;           // host threads staging a pageable input through the pinned ring; 0 = let the
                                   // driver copy from pageable memory itself
  int cold_staged = 1;             // first host-buffer call on a ctx returns its CSR in unpinned memory filled
                                   // through a pinned ring (0: pin the result arrays at once, as later calls do)
  int optimistic = 1;              // launch a pass's kernels without waiting for the plan's totals when the
                                   // previous pass on this ctx left sizes to go by (one host sync instead of three)
  int epilogue = 0;                // 1: sort + CSR write inside the fused kernel instead of the count scan +
                                   // k_sort_write.  Measured slower (cfg4: 4.04 ms against 2.69 + 0.80): at the
                                   // same 32 warps per SM a warp's epilogue is just appended to its serial
                                   // chain, and warps parked in its latency-bound loops leave fewer warps to
                                   // fill the issue slots (64.6 % issue-active against 77 %; profiles/r02f)
};

// Pageable host memory backed by transparent huge pages (mmap + MADV_HUGEPAGE): what the FIRST host-buffer call
// on a ctx returns its CSR in.  Pinning 1.1 GB takes 0.55-1.3 s on this pool's hosts
// (profiles/r02b_alloc_probe.txt, r02l_cold_probe.json) — far longer than computing and transferring the result —
// so a one-shot caller gets an unpinned array filled through a small pinned ring by host threads (Drainer);
// a second call on the same ctx is no one-shot any more and switches to pinned arrays.
struct PageBuf {
  void *p = nullptr;
  size_t cap = 0;
  int ensure_fresh(size_t bytes) {  // contents are not kept
    if (bytes <= cap) return C2B_OK;
    release();
    const size_t want = (bytes + bytes / 16 + (2u << 20)) & ~(size_t)((2u << 20) - 1);
#if defined(__linux__)
    void *q = mmap(nullptr, want, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (q == MAP_FAILED) return set_error(C2B_ERR_OOM, "mmap(%zu bytes) failed", want);
    (void)madvise(q, want, MADV_HUGEPAGE);
#else
    void *q = malloc(want);
    if (!q) return set_error(C2B_ERR_OOM, "malloc(%zu bytes) failed", want);
#endif
    p = q;
    cap = want;
    return C2B_OK;
  }
  void release() {
#if defined(__linux__)
    if (p) munmap(p, cap);
#else
    free(p);
#endif
    p = nullptr;
    cap = 0;
  }
  template <class T>
  T *as() const {
    return reinterpret_cast<T *>(p);
  }
};

// Device -> pageable host memory through a pinned ring: `threads` host threads, each with its own stream and two
// 2 MB ring slots, take 2 MB chunks off a queue: DMA into a slot, then memcpy the slot to its destination (the
// first touch of the destination's pages happens there, in parallel).  submit() never blocks the caller.
class Drainer {
 public:
  static constexpr size_t CH = 2u << 20;  // pinning the ring itself costs ~1 ms per MB: keep it small
  int start(int device, int threads, char *ring) {
    n_ = threads;
    for (int t = 0; t < n_; ++t)
      workers_.emplace_back([this, device, t, ring]() { loop(device, ring + (size_t)t * 2 * CH); });
    return C2B_OK;
  }
  // bytes from dev_src (valid once `ready` has fired on its stream) to host_dst; `batch` tags the chunks
  void submit(const void *dev_src, void *host_dst, size_t bytes, cudaEvent_t ready, int batch) {
    std::lock_guard<std::mutex> lk(mu_);
    if ((size_t)batch >= dma_left_.size()) dma_left_.resize((size_t)batch + 1, 0);
    for (size_t off = 0; off < bytes; off += CH) {
      q_.push_back(Chunk{static_cast<const char *>(dev_src) + off, static_cast<char *>(host_dst) + off,
                         std::min(CH, bytes - off), ready, batch});
      ++dma_left_[(size_t)batch];
      ++left_;
    }
    cv_.notify_all();
  }
  // every chunk of `batch` has left the device (its device buffers may be overwritten)
  void wait_dma(int batch) {
    std::unique_lock<std::mutex> lk(mu_);
    cv_done_.wait(lk, [&] { return (size_t)batch >= dma_left_.size() || dma_left_[(size_t)batch] == 0 || err_ != cudaSuccess; });
  }
  // everything submitted so far has landed in host memory
  void wait_all() {
    std::unique_lock<std::mutex> lk(mu_);
    cv_done_.wait(lk, [&] { return left_ == 0 || err_ != cudaSuccess; });
  }
  // everything submitted has landed in host memory; joins the threads
  cudaError_t finish() {
    {
      std::unique_lock<std::mutex> lk(mu_);
      cv_done_.wait(lk, [&] { return left_ == 0 || err_ != cudaSuccess; });
      quit_ = true;
    }
    cv_.notify_all();
    for (auto &w : workers_) w.join();
    workers_.clear();
    return err_;
  }
  ~Drainer() {
    if (!workers_.empty()) (void)finish();
  }

 private:
  struct Chunk {
    const char *src;
    char *dst;
    size_t n;
    cudaEvent_t ready;
    int batch;
  };
  void loop(int device, char *slots) {
    cudaError_t e = cudaSetDevice(device);
    cudaStream_t st = nullptr;
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    for (int turn = 0;; ++turn) {
      Chunk c;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return !q_.empty() || quit_ || err_ != cudaSuccess; });
        if (q_.empty() || err_ != cudaSuccess) break;
        c = q_.front();
        q_.pop_front();
      }
      char *slot = slots + (size_t)(turn & 1) * CH;
      if (e == cudaSuccess) e = cudaStreamWaitEvent(st, c.ready, 0);
      if (e == cudaSuccess) e = cudaMemcpyAsync(slot, c.src, c.n, cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
      {
        std::lock_guard<std::mutex> lk(mu_);
        if (e != cudaSuccess) err_ = e;
        --dma_left_[(size_t)c.batch];
      }
      cv_done_.notify_all();
      if (e == cudaSuccess) memcpy(c.dst, slot, c.n);
      {
        std::lock_guard<std::mutex> lk(mu_);
        --left_;
      }
      cv_done_.notify_all();
      if (e != cudaSuccess) break;
    }
    if (st) cudaStreamDestroy(st);
    cv_done_.notify_all();
  }
  int n_ = 0;
  std::vector<std::thread> workers_;
  std::mutex mu_;
  std::condition_variable cv_, cv_done_;
  std::deque<Chunk> q_;
  std::vector<size_t> dma_left_;
  size_t left_ = 0;
  bool quit_ = false;
  cudaError_t err_ = cudaSuccess;
};

// what the previous host-buffer call cost per camera: decides whether the next one is bound by the result
// transfer or by the kernels (camera batch policy of c2b_visibility_graph)
struct CallHist {
  bool valid = false;
  double ms_compute_per_cam = 0, d2h_bytes_per_cam = 0;
};

// what the previous resident pass (grid schedule) found: lets the next one launch its kernels without waiting
// for the plan's totals (the kernels check every assumption and flag a misfit)
struct PassMemo {
  bool valid = false, any_overflow = false, big = false;
};

struct CtxExtra {
  Tunables tun;
  CallHist hist;
  PassMemo memo;
  bool grid_locked = false;  // inside c2b_visibility_graph: the grid was built for all of the call's cameras
  PageBuf pg_idx, pg_uv;     // first host-buffer call: unpinned result arrays (see PageBuf)
  PinBuf pin_out;            // and the ring they are filled through
  ~CtxExtra() {
    pg_idx.release();
    pg_uv.release();
    pin_out.release();
  }
  GridCache grid;
  uint64_t points_version = 0;
  double pts_bounds[6] = {0, 0, 0, 0, 0, 0};
  bool have_points = false, have_cameras = false, have_result = false;
  uint64_t res_candidates = 0;
};

}  // namespace

// c2b_ctx carries its non-POD extras behind one pointer so the struct in c2b_common.cuh stays plain
static std::vector<std::pair<c2b_ctx *, CtxExtra *>> &extras() {
  static std::vector<std::pair<c2b_ctx *, CtxExtra *>> v;
  return v;
}
static CtxExtra *extra_of(c2b_ctx *ctx) {
  for (auto &e : extras())
    if (e.first == ctx) return e.second;
  return nullptr;
}

static bool tune(Tunables &t, const char *name, double v) {
  const std::string n(name);
  if (n == "parts_log2") t.parts_log2 = v < 0 ? -1 : std::min((int)v, 2);
  else if (n == "trilist_cap") t.trilist_cap = std::max(1, (int)v);
  else if (n == "max_pairs") t.max_pairs = std::min<uint64_t>(0xffffffffull, v < 0 ? 0 : (uint64_t)v);
  else if (n == "hoist_max") t.hoist_max = std::min(std::max(0, (int)v), 64);
  else if (n == "packet_bvh") t.packet_bvh = v != 0;
  else if (n == "trilist_warp") t.trilist_warp = v < 0 ? -1 : (v != 0);
  else if (n == "batches") t.batches = std::max(0, (int)v);
  else if (n == "fu_occ3") t.fu_occ3 = v != 0;
  else if (n == "grid_cell_factor") t.grid_cell_factor = v > 0 ? v : 0.25;
  else if (n == "stage_threads") t.stage_threads = std::min(std::max(0, (int)v), 16);
  else if (n == "epilogue") t.epilogue = v != 0;
  else if (n == "optimistic") t.optimistic = v != 0;
  else if (n == "cold_staged") t.cold_staged = v != 0;
  else return false;
  return true;
}

static float scene_abs_max(const c2b_scene *scene) {
  float m = 0.0f;
  if (scene)
    for (int k = 0; k < 3; ++k) m = std::max(m, std::max(std::fabs(scene->lo[k]), std::fabs(scene->hi[k])));
  if (!std::isfinite(m)) m = 3.0e38f;
  return m;
}


extern "C" {

const char *c2b_last_error(void) { return last_error_ref().c_str(); }
int c2b_abi_version(void) { return C2B_ABI_VERSION; }
uint64_t c2b_kernel_launches(void) { return launch_counter(); }

int c2b_init(int device, c2b_ctx **out) {
  if (!out) return set_error(C2B_ERR_INVALID, "c2b_init: out is null");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    (void)cudaGetLastError();
    return set_error(C2B_ERR_NO_DEVICE, "no CUDA device visible (%s); this library has no CPU fallback",
                     e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  }
  if (device < 0 || device >= n) return set_error(C2B_ERR_INVALID, "device %d out of range [0,%d)", device, n);
  C2B_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  C2B_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    return set_error(C2B_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only",
                     device, prop.major, prop.minor);
  c2b_ctx *ctx = new (std::nothrow) c2b_ctx();
  if (!ctx) return set_error(C2B_ERR_OOM, "out of host memory");
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  C2B_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  for (int i = 0; i < EV_COUNT; ++i) C2B_CUDA(cudaEventCreate(&ctx->ev[i]));
  C2B_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  C2B_CUDA(cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking));
  C2B_CUDA(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
  C2B_CUDA(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
  for (int i = 0; i < 2; ++i) {
    C2B_CUDA(cudaEventCreateWithFlags(&ctx->ev_ready[i], cudaEventDisableTiming));
    C2B_CUDA(cudaEventCreateWithFlags(&ctx->ev_copied[i], cudaEventDisableTiming));
  }
  // NUMA node of the device (sysfs), for the pinned staging / result buffers (see PinBuf)
  {
    int node = -1;
    const char *off = getenv("C2B_NUMA_LOCAL");
    char bus[32] = {0};
    if (!(off && atoi(off) == 0) && cudaDeviceGetPCIBusId(bus, sizeof bus, device) == cudaSuccess) {
      for (char *c = bus; *c; ++c) *c = (char)tolower((unsigned char)*c);
      const std::string path = std::string("/sys/bus/pci/devices/") + bus + "/numa_node";
      if (FILE *f = fopen(path.c_str(), "r")) {
        if (fscanf(f, "%d", &node) != 1) node = -1;
        fclose(f);
      }
    }
    (void)cudaGetLastError();
    ctx->numa_node = node;
    for (PinBuf *b : {&ctx->pin_in, &ctx->h_offsets, &ctx->h_idx, &ctx->h_uv, &ctx->h_small}) b->node = node;
  }
  {  // tables of the noise draws (c2b_noise.cuh), one copy per device
    static double2 sc[256], ln[128];
    static std::once_flag once;
    std::call_once(once, [] { noise_tables_host(sc, ln); });
    C2B_CUDA(cudaMemcpyToSymbol(g_sc_tab, sc, sizeof sc));
    C2B_CUDA(cudaMemcpyToSymbol(g_ln_tab, ln, sizeof ln));
  }
  CtxExtra *x = new CtxExtra();
  {
    struct { const char *env, *name; } hooks[] = {
        {"C2B_PARTS_LOG2", "parts_log2"}, {"C2B_TRILIST_CAP", "trilist_cap"}, {"C2B_MAX_PAIRS", "max_pairs"},
        {"C2B_HOIST_MAX", "hoist_max"}, {"C2B_PACKET_BVH", "packet_bvh"}, {"C2B_TRILIST_WARP", "trilist_warp"},
        {"C2B_BATCHES", "batches"}, {"C2B_FU_OCC3", "fu_occ3"}, {"C2B_GRID_CELL_FACTOR", "grid_cell_factor"},
        {"C2B_STAGE_THREADS", "stage_threads"}, {"C2B_EPILOGUE", "epilogue"}, {"C2B_OPTIMISTIC", "optimistic"}, {"C2B_COLD_STAGED", "cold_staged"}};
    for (auto &h : hooks)
      if (const char *e = getenv(h.env)) (void)tune(x->tun, h.name, atof(e));
  }
  extras().push_back({ctx, x});
  *out = ctx;
  return C2B_OK;
}

int c2b_tune(c2b_ctx *ctx, const char *name, double value) {
  if (!ctx || !name) return set_error(C2B_ERR_INVALID, "c2b_tune: null argument");
  if (!strcmp(name, "reset")) {
    extra_of(ctx)->tun = Tunables();
    return C2B_OK;
  }
  if (!tune(extra_of(ctx)->tun, name, value)) return set_error(C2B_ERR_INVALID, "c2b_tune: unknown hook '%s'", name);
  return C2B_OK;
}

void c2b_shutdown(c2b_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
  if (ctx->aux_stream) cudaStreamSynchronize(ctx->aux_stream);
  DevBuf *bufs[] = {&ctx->cams, &ctx->cam_center, &ctx->pts, &ctx->stage, &ctx->cell_of_pt,
                    &ctx->cell_start, &ctx->cell_cursor, &ctx->grid_x, &ctx->grid_y, &ctx->grid_z,
                    &ctx->grid_idx, &ctx->pool_key, &ctx->pool_uv, &ctx->cam_count, &ctx->counters,
                    &ctx->sort_keys[0], &ctx->sort_keys[1], &ctx->sort_vals[0], &ctx->sort_vals[1],
                    &ctx->sort_hist, &ctx->scan_tmp[0], &ctx->scan_tmp[1], &ctx->scan_tmp[2],
                    &ctx->vis_words, &ctx->word_prefix, &ctx->out_offsets[0], &ctx->out_idx[0],
                    &ctx->out_uv[0], &ctx->out_offsets[1], &ctx->out_idx[1], &ctx->out_uv[1],
                    &ctx->misc, &ctx->tri_list, &ctx->tri_count, &ctx->pts_aos, &ctx->ev_off,
                    &ctx->vis_count, &ctx->seg_off, &ctx->scratch_idx, &ctx->plan_rows, &ctx->plan_row_count,
                    &ctx->nz_cams, &ctx->nz_centers, &ctx->nz_pts, &ctx->nz_uv, &ctx->nz_scratch,
                    &ctx->epi_status, &ctx->epi_prefix, &ctx->cams_all};
  for (auto *b : bufs) b->release();
  PinBuf *pins[] = {&ctx->pin_in, &ctx->h_offsets, &ctx->h_idx, &ctx->h_uv, &ctx->h_small};
  for (auto *p : pins) p->release();
  for (int i = 0; i < EV_COUNT; ++i)
    if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
  for (int i = 0; i < 2; ++i) {
    if (ctx->ev_ready[i]) cudaEventDestroy(ctx->ev_ready[i]);
    if (ctx->ev_copied[i]) cudaEventDestroy(ctx->ev_copied[i]);
  }
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->aux_stream) cudaStreamDestroy(ctx->aux_stream);
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  auto &v = extras();
  for (size_t i = 0; i < v.size(); ++i)
    if (v[i].first == ctx) {
      delete v[i].second;
      v.erase(v.begin() + i);
      break;
    }
  delete ctx;
}

// ---- scene -------------------------------------------------------------------------------------
int c2b_scene_create(c2b_ctx *ctx, const float *xyz, uint64_t nv, const uint32_t *tri, uint64_t nt,
                     c2b_scene **out) {
  if (!ctx || !out) return set_error(C2B_ERR_INVALID, "c2b_scene_create: null ctx/out");
  *out = nullptr;
  if ((nv && !xyz) || (nt && !tri)) return set_error(C2B_ERR_INVALID, "c2b_scene_create: null mesh arrays");
  C2B_CUDA(cudaSetDevice(ctx->device));
  std::vector<uint32_t> keep;
  keep.reserve(3 * nt);
  for (uint64_t i = 0; i < nt; ++i) {
    uint32_t a = tri[3 * i], b = tri[3 * i + 1], c = tri[3 * i + 2];
    if (a >= nv || b >= nv || c >= nv)
      return set_error(C2B_ERR_INVALID, "triangle %llu references vertex >= %llu",
                       (unsigned long long)i, (unsigned long long)nv);
    if (a == b || b == c || a == c) continue;  // zero-area (tobj `l` records): can never be hit
    keep.push_back(a);
    keep.push_back(b);
    keep.push_back(c);
  }
  c2b_scene *sc = new (std::nothrow) c2b_scene();
  if (!sc) return set_error(C2B_ERR_OOM, "out of host memory");
  sc->device = ctx->device;
  int rc = bvh_build(ctx, sc, xyz, nv, keep.data(), keep.size() / 3);
  if (rc != C2B_OK) {
    sc->nodes.release();
    sc->tris.release();
    delete sc;
    return rc;
  }
  *out = sc;
  return C2B_OK;
}

int c2b_scene_bounds(const c2b_scene *scene, float lower[3], float upper[3]) {
  if (!scene || !lower || !upper) return set_error(C2B_ERR_INVALID, "c2b_scene_bounds: null argument");
  for (int k = 0; k < 3; ++k) {
    lower[k] = scene->lo[k];
    upper[k] = scene->hi[k];
  }
  return C2B_OK;
}
uint64_t c2b_scene_num_triangles(const c2b_scene *scene) { return scene ? scene->n_tris : 0; }
uint64_t c2b_scene_num_nodes(const c2b_scene *scene) { return scene ? scene->n_nodes : 0; }
void c2b_scene_destroy(c2b_scene *scene) {
  if (!scene) return;
  cudaSetDevice(scene->device);
  scene->nodes.release();
  scene->tris.release();
  delete scene;
}

// ---- ray-level entries ------------------------------------------------------------------------------
int c2b_occluded(c2b_ctx *ctx, const c2b_scene *scene, c2b_ray48 *rays, uint64_t n) {
  if (!ctx || !scene || (n && !rays)) return set_error(C2B_ERR_INVALID, "c2b_occluded: null argument");
  if (n == 0 || scene->n_nodes == 0) return C2B_OK;
  C2B_CUDA(cudaSetDevice(ctx->device));
  C2B_TRY(ctx->stage.ensure(n * sizeof(c2b_ray48)));
  C2B_TRY(ctx->counters.ensure(32));
  float absmax = 0.0f;
  for (int k = 0; k < 3; ++k) absmax = std::max(absmax, std::max(std::fabs(scene->lo[k]), std::fabs(scene->hi[k])));
  C2B_CUDA(cudaMemcpyAsync(ctx->stage.p, rays, n * sizeof(c2b_ray48), cudaMemcpyHostToDevice, ctx->stream));
  k_occluded_rays<false><<<blocks_for(n, 256), 256, 0, ctx->stream>>>(
      scene->nodes.as<float4>(), scene->tris.as<float4>(), (int)scene->n_nodes,
      ctx->stage.as<c2b_ray48>(), n, absmax, ctx->counters.as<unsigned long long>());
  C2B_KERNEL_CHECK();
  C2B_CUDA(cudaMemcpyAsync(rays, ctx->stage.p, n * sizeof(c2b_ray48), cudaMemcpyDeviceToHost, ctx->stream));
  C2B_CUDA(cudaStreamSynchronize(ctx->stream));
  return C2B_OK;
}

int c2b_intersect(c2b_ctx *ctx, const c2b_scene *scene, c2b_ray48 *rays, uint64_t n) {
  if (!ctx || !scene || (n && !rays)) return set_error(C2B_ERR_INVALID, "c2b_intersect: null argument");
  if (n == 0) return C2B_OK;
  if (scene->n_nodes == 0) {
    for (uint64_t i = 0; i < n; ++i) rays[i].flags = 0;
    return C2B_OK;
  }
  C2B_CUDA(cudaSetDevice(ctx->device));
  C2B_TRY(ctx->stage.ensure(n * sizeof(c2b_ray48)));
  C2B_CUDA(cudaMemcpyAsync(ctx->stage.p, rays, n * sizeof(c2b_ray48), cudaMemcpyHostToDevice, ctx->stream));
  k_intersect_rays<<<blocks_for(n, 256), 256, 0, ctx->stream>>>(scene->nodes.as<float4>(), scene->tris.as<float4>(),
                                                                (int)scene->n_nodes, ctx->stage.as<c2b_ray48>(), n,
                                                                scene_abs_max(scene));
  C2B_KERNEL_CHECK();
  C2B_CUDA(cudaMemcpyAsync(rays, ctx->stage.p, n * sizeof(c2b_ray48), cudaMemcpyDeviceToHost, ctx->stream));
  C2B_CUDA(cudaStreamSynchronize(ctx->stream));
  return C2B_OK;
}

int c2b_intersect1(c2b_ctx *ctx, const c2b_scene *scene, const float org[3], const float dir[3],
                   int *hit, float *tfar) {
  if (!ctx || !scene || !org || !dir || !hit || !tfar)
    return set_error(C2B_ERR_INVALID, "c2b_intersect1: null argument");
  c2b_ray48 r;
  memset(&r, 0, sizeof r);
  r.org_x = org[0];
  r.org_y = org[1];
  r.org_z = org[2];
  r.dir_x = dir[0];
  r.dir_y = dir[1];
  r.dir_z = dir[2];
  r.tfar = INFINITY;  // Ray::new(org, dir): tnear 0, tfar +inf
  C2B_TRY(c2b_intersect(ctx, scene, &r, 1));
  *hit = r.flags != 0;
  *tfar = r.flags ? r.tfar : INFINITY;
  return C2B_OK;
}

// ---- uploads ------------------------------------------------------------------------------------------
void c2b_vis_options_default(c2b_vis_options *opt) {
  if (!opt) return;
  opt->cull_mode = C2B_CULL_GRID;
  opt->occlusion = C2B_OCC_MESH;
  opt->endpoint_guard_rel = 0;
  opt->count_traversal = 0;
  opt->block_length = 20.0;
  opt->block_inset = 1.0;
  opt->predicate = C2B_PRED_WATERTIGHT;
  opt->reserved = 0;
}

// Host -> device copy of a caller's array on ctx->stream.  A pinned (or registered) source goes to the copy
// engine as it is.  A PAGEABLE one — what `points.as_ptr()` of a Rust Vec or a numpy array is — would be
// staged by the driver through its own bounce buffer on one thread; instead `stage_threads` host threads
// copy 2 MB chunks into the ctx's pinned ring (two slots per thread) and queue each chunk's DMA as soon as
// it is staged, so the host-side memcpy runs at several cores' bandwidth and overlaps the PCIe transfer.
int copy_in(c2b_ctx *ctx, void *d, const void *h, size_t bytes, cudaMemcpyKind kind) {
  cudaStream_t st = ctx->stream;
  if (bytes == 0) return C2B_OK;
  const int T = extra_of(ctx)->tun.stage_threads;
  constexpr size_t CH = 2u << 20;  // pinning the ring itself costs ~1 ms per MB: keep it small
  bool pageable = false;
  if (kind == cudaMemcpyHostToDevice && T > 0 && bytes >= (64u << 10)) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, h) == cudaSuccess)
      pageable = at.type == cudaMemoryTypeUnregistered;
    else
      (void)cudaGetLastError();
  }
  if (!pageable) {
    C2B_CUDA(cudaMemcpyAsync(d, h, bytes, kind, st));
    return C2B_OK;
  }
  if (bytes < 4 * CH) {
    // a small pageable array (a camera range): one hop through the ring on this thread.  Left to the driver,
    // pageable copies issued for several GPUs at once queue behind one another — 12 ms for 1.5 MB per GPU at
    // N = 8 (profiles/r02k_bench_cfg4_8gpu.json: e2e_pageable)
    C2B_CUDA(cudaStreamSynchronize(st));  // the ring may still feed an earlier copy
    C2B_TRY(ctx->pin_in.ensure(std::max<size_t>(bytes, 8 * CH)));
    memcpy(ctx->pin_in.p, h, bytes);
    C2B_CUDA(cudaMemcpyAsync(d, ctx->pin_in.p, bytes, cudaMemcpyHostToDevice, st));
    return C2B_OK;
  }
  const size_t n_chunks = (bytes + CH - 1) / CH;
  const int nt = (int)std::min<size_t>((size_t)T, n_chunks);
  // the ring may still feed the previous call's DMA
  C2B_CUDA(cudaStreamSynchronize(st));
  C2B_TRY(ctx->pin_in.ensure(std::max<size_t>((size_t)nt * 2 * CH, 8 * CH)));
  std::vector<cudaError_t> err((size_t)nt, cudaSuccess);
  std::vector<std::thread> workers;
  const int device = ctx->device;
  char *ring = ctx->pin_in.as<char>();
  for (int t = 0; t < nt; ++t)
    workers.emplace_back([=, &err]() {
      cudaError_t e = cudaSetDevice(device);
      cudaEvent_t ev[2] = {nullptr, nullptr};
      for (int k = 0; k < 2 && e == cudaSuccess; ++k) e = cudaEventCreateWithFlags(&ev[k], cudaEventDisableTiming);
      size_t turn = 0;
      for (size_t c = (size_t)t; c < n_chunks && e == cudaSuccess; c += (size_t)nt, ++turn) {
        const int slot = (int)(turn & 1);
        char *pin = ring + ((size_t)t * 2 + slot) * CH;
        const size_t off = c * CH, n = std::min(CH, bytes - off);
        if (turn >= 2) e = cudaEventSynchronize(ev[slot]);  // the slot's previous DMA has drained
        if (e != cudaSuccess) break;
        memcpy(pin, static_cast<const char *>(h) + off, n);
        e = cudaMemcpyAsync(static_cast<char *>(d) + off, pin, n, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaEventRecord(ev[slot], st);
      }
      for (int k = 0; k < 2; ++k)
        if (ev[k]) {
          if (e == cudaSuccess) e = cudaEventSynchronize(ev[k]);
          cudaEventDestroy(ev[k]);
        }
      err[(size_t)t] = e;
    });
  for (auto &w : workers) w.join();
  for (cudaError_t e : err)
    if (e != cudaSuccess) {
      (void)cudaGetLastError();
      return set_error(C2B_ERR_CUDA, "staged host-to-device copy failed: %s", cudaGetErrorString(e));
    }
  return C2B_OK;
}

int c2b_internal_copy_in(c2b_ctx *ctx, void *d, const void *h, size_t bytes) {
  return copy_in(ctx, d, h, bytes, cudaMemcpyHostToDevice);
}

// The AoS point array of the ctx, with room for `capacity` points, for callers that fill it on the device
// themselves (an NCCL all-gather of per-GPU shards, c2b_multi.cu); followed by c2b_points_commit.
int c2b_points_device_buffer(c2b_ctx *ctx, uint64_t capacity, double **d_out) {
  if (!ctx || !d_out) return set_error(C2B_ERR_INVALID, "c2b_points_device_buffer: null argument");
  C2B_CUDA(cudaSetDevice(ctx->device));
  C2B_TRY(ctx->pts_aos.ensure(std::max<uint64_t>(capacity, 1) * 24));
  *d_out = ctx->pts_aos.as<double>();
  return C2B_OK;
}

// pts_aos[0, P) is (being) written on ctx->stream: derive the SoA copy and the coordinate bounds
int c2b_points_commit(c2b_ctx *ctx, uint64_t P) {
  if (!ctx) return set_error(C2B_ERR_INVALID, "c2b_points_commit: null ctx");
  if (P * 24 > ctx->pts_aos.cap) return set_error(C2B_ERR_INVALID, "c2b_points_commit: %llu points exceed the buffer", (unsigned long long)P);
  if (P >= 0xffffffffull) return set_error(C2B_ERR_INVALID, "too many points (%llu)", (unsigned long long)P);
  CtxExtra *x = extra_of(ctx);
  C2B_CUDA(cudaSetDevice(ctx->device));
  ctx->P = P;
  x->points_version++;
  x->grid.valid = false;
  x->have_points = true;
  x->have_result = false;
  if (P == 0) return C2B_OK;
  C2B_TRY(ctx->pts.ensure(P * 24));
  double *px = ctx->pts.as<double>(), *py = px + P, *pz = py + P;
  k_aos_to_soa3<<<blocks_for(P, 256), 256, 0, ctx->stream>>>(ctx->pts_aos.as<double>(), P, px, py, pz);
  C2B_KERNEL_CHECK();
  // coordinate bounds (grid extent); tiny D2H happens lazily when a grid is built
  int nb = (int)std::min<uint64_t>(blocks_for(P, 256), (uint64_t)4 * ctx->sm_count);
  C2B_TRY(ctx->misc.ensure((size_t)nb * 48 + 64));
  double *partial = ctx->misc.as<double>() + 8;
  k_pts_bounds_partial<<<nb, 256, 0, ctx->stream>>>(px, py, pz, P, partial);
  C2B_KERNEL_CHECK();
  k_pts_bounds_final<<<1, 32, 0, ctx->stream>>>(partial, nb, ctx->misc.as<double>());
  C2B_KERNEL_CHECK();
  C2B_CUDA(cudaMemcpyAsync(x->pts_bounds, ctx->misc.p, 48, cudaMemcpyDeviceToHost, ctx->stream));
  return C2B_OK;
}

static int upload_points_impl(c2b_ctx *ctx, const double *pts, uint64_t P, cudaMemcpyKind kind) {
  if (!ctx || (P && !pts)) return set_error(C2B_ERR_INVALID, "c2b_upload_points: null argument");
  if (P >= 0xffffffffull) return set_error(C2B_ERR_INVALID, "too many points (%llu)", (unsigned long long)P);
  double *d = nullptr;
  C2B_TRY(c2b_points_device_buffer(ctx, P, &d));
  if (P) C2B_TRY(copy_in(ctx, d, pts, P * 24, kind));
  return c2b_points_commit(ctx, P);
}

int c2b_upload_points(c2b_ctx *ctx, const double *pts, uint64_t P) {
  return upload_points_impl(ctx, pts, P, cudaMemcpyHostToDevice);
}

int c2b_upload_points_device(c2b_ctx *ctx, const double *d_pts, uint64_t P) {
  return upload_points_impl(ctx, d_pts, P, cudaMemcpyDeviceToDevice);
}

int c2b_drop_point_grid(c2b_ctx *ctx) {
  if (!ctx) return set_error(C2B_ERR_INVALID, "c2b_drop_point_grid: null ctx");
  extra_of(ctx)->grid.valid = false;
  return C2B_OK;
}

int c2b_upload_cameras(c2b_ctx *ctx, const double *cams, uint64_t C) {
  if (!ctx || (C && !cams)) return set_error(C2B_ERR_INVALID, "c2b_upload_cameras: null argument");
  if (C >= 0xffffffffull) return set_error(C2B_ERR_INVALID, "too many cameras (%llu)", (unsigned long long)C);
  CtxExtra *x = extra_of(ctx);
  C2B_CUDA(cudaSetDevice(ctx->device));
  ctx->C = C;
  x->have_cameras = true;
  x->have_result = false;
  if (C == 0) return C2B_OK;
  C2B_TRY(ctx->cams.ensure(C * 120));
  C2B_TRY(ctx->cam_center.ensure(C * 24));
  C2B_TRY(copy_in(ctx, ctx->cams.p, cams, C * 120, cudaMemcpyHostToDevice));
  double *cx = ctx->cam_center.as<double>();
  k_cam_prep<<<blocks_for(C, 128), 128, 0, ctx->stream>>>(ctx->cams.as<double>(), C, cx, cx + C, cx + 2 * C);
  C2B_KERNEL_CHECK();
  return C2B_OK;
}

// cameras [first, first + count) of the call's device copy (ctx->cams_all) become the resident cameras:
// what c2b_upload_cameras does for a batch, without touching host memory again
static int select_cameras(c2b_ctx *ctx, uint64_t first, uint64_t count) {
  CtxExtra *x = extra_of(ctx);
  ctx->C = count;
  x->have_cameras = true;
  x->have_result = false;
  if (count == 0) return C2B_OK;
  C2B_TRY(ctx->cams.ensure(count * 120));
  C2B_TRY(ctx->cam_center.ensure(count * 24));
  C2B_CUDA(cudaMemcpyAsync(ctx->cams.p, ctx->cams_all.as<double>() + 15 * first, count * 120, cudaMemcpyDeviceToDevice,
                           ctx->stream));
  double *cx = ctx->cam_center.as<double>();
  k_cam_prep<<<blocks_for(count, 128), 128, 0, ctx->stream>>>(ctx->cams.as<double>(), count, cx, cx + count, cx + 2 * count);
  C2B_KERNEL_CHECK();
  return C2B_OK;
}

// ---- the pipeline ---------------------------------------------------------------------------------------
// the 128-byte counter block -> host, synchronising the compute stream
static int read_counters(c2b_ctx *ctx, unsigned long long *h_cnt) {
  C2B_TRY(ctx->h_small.ensure(256));
  k_copy_words<<<1, 32, 0, ctx->stream>>>(ctx->counters.as<uint32_t>(), ctx->h_small.as<uint32_t>(), 32);
  C2B_KERNEL_CHECK();
  C2B_CUDA(cudaStreamSynchronize(ctx->stream));
  memcpy(h_cnt, ctx->h_small.p, 128);
  return C2B_OK;
}

// bounding box of the resident cameras' centres (tiny two-stage reduction; synchronises the stream)
static int camera_bbox(c2b_ctx *ctx, double lo[3], double hi[3]) {
  const uint64_t C = ctx->C;
  const double *cx = ctx->cam_center.as<double>();
  int nb = (int)std::min<uint64_t>(std::max<uint64_t>(blocks_for(C, 256), 1), (uint64_t)ctx->sm_count);
  C2B_TRY(ctx->misc.ensure((size_t)nb * 48 + 64));
  double *partial = ctx->misc.as<double>() + 8;
  k_pts_bounds_partial<<<nb, 256, 0, ctx->stream>>>(cx, cx + C, cx + 2 * C, C, partial);
  C2B_KERNEL_CHECK();
  k_pts_bounds_final<<<1, 32, 0, ctx->stream>>>(partial, nb, ctx->misc.as<double>());
  C2B_KERNEL_CHECK();
  double b[6];
  C2B_CUDA(cudaMemcpyAsync(b, ctx->misc.p, 48, cudaMemcpyDeviceToHost, ctx->stream));
  C2B_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int k = 0; k < 3; ++k) {
    lo[k] = b[k];
    hi[k] = b[3 + k];
  }
  return C2B_OK;
}

// what the cameras can reach: their centres' box grown by max_dist (and a little)
static void reach_box(const double clo[3], const double chi[3], double max_dist, double lo[3], double hi[3]) {
  for (int k = 0; k < 3; ++k) {
    const double pad = max_dist + 1e-6 * (std::fabs(max_dist) + std::fabs(clo[k]) + std::fabs(chi[k]));
    lo[k] = clo[k] - pad;
    hi[k] = chi[k] + pad;
  }
}

static int build_grid(c2b_ctx *ctx, CtxExtra *x, double max_dist) {
  cudaStream_t st = ctx->stream;
  const uint64_t P = ctx->P;
  C2B_CUDA(cudaStreamSynchronize(st));  // pts_bounds has landed
  GridDesc g;
  double h = max_dist * x->tun.grid_cell_factor;
  double ext[3];
  for (int k = 0; k < 3; ++k) {
    g.min_c[k] = x->pts_bounds[k];
    g.max_c[k] = x->pts_bounds[3 + k];
    ext[k] = g.max_c[k] - g.min_c[k];
    if (!(ext[k] >= 0.0) || std::isinf(ext[k])) ext[k] = 0.0;  // NaN / inf coordinates: one cell
    g.lo[k] = std::isfinite(g.min_c[k]) ? g.min_c[k] : 0.0;
  }
  if (!(h > 0.0) || !std::isfinite(h)) h = 1.0;
  const double max_cells = 4194304.0;  // 2^22
  auto cells_at = [&](double hh) {
    double total = 1.0;
    for (int k = 0; k < 3; ++k) total *= std::min(std::floor(ext[k] / hh) + 1.0, 1e9);
    return total;
  };
  x->grid.hinted = false;
  g.drop = 0;
  for (int k = 0; k < 3; ++k) g.rlo[k] = g.rhi[k] = 0.0;
  bool odd_bounds = false;  // an infinite coordinate would collapse its axis to one cell
  for (int k = 0; k < 3; ++k) odd_bounds = odd_bounds || !std::isfinite(g.min_c[k]) || !std::isfinite(g.max_c[k]);
  if (ctx->C) {
    // No point farther than max_dist from every camera can be observed, so only the cameras' REACH (their
    // centres' box grown by max_dist) has to be binned.  When the reach is clearly smaller than the data —
    // a camera shard of a multi-GPU run, or a few far-away outlier vertices of an OBJ scene that would stretch
    // every cell — the cells cover the reach only and points outside it are left out of the grid altogether
    // (k_grid_count); a GPU that owns 1/8 of the cameras then bins about 1/8 of the points.
    double clo[3], chi[3], rlo[3], rhi[3];
    C2B_TRY(camera_bbox(ctx, clo, chi));
    reach_box(clo, chi, max_dist, rlo, rhi);
    bool usable = true;
    for (int k = 0; k < 3; ++k) usable = usable && std::isfinite(rlo[k]) && std::isfinite(rhi[k]) && rlo[k] <= rhi[k];
    double frac = 1.0;  // share of the data's box the reach covers
    if (usable)
      for (int k = 0; k < 3; ++k) {
        const double a = std::fmax(g.min_c[k], rlo[k]), b = std::fmin(g.max_c[k], rhi[k]);
        if (ext[k] > 0.0) frac *= a <= b ? std::fmin((b - a) / ext[k], 1.0) : 0.0;
      }
    if (usable && (cells_at(h) > max_cells || odd_bounds || frac < 0.7)) {
      for (int k = 0; k < 3; ++k) {
        // the reach intersected with the data bounds (which may be infinite); empty: one cell at the reach's edge
        const double a = std::fmax(g.min_c[k], rlo[k]), b = std::fmin(g.max_c[k], rhi[k]);
        g.lo[k] = a <= b ? a : rlo[k];
        ext[k] = a <= b ? b - a : 0.0;
        g.min_c[k] = g.lo[k];  // what is IN the grid: rows of edge cells end here (camera_row)
        g.max_c[k] = a <= b ? b : g.lo[k];
        g.rlo[k] = rlo[k];
        g.rhi[k] = rhi[k];
        x->grid.hint_lo[k] = rlo[k];
        x->grid.hint_hi[k] = rhi[k];
      }
      g.drop = 1;
      x->grid.hinted = true;
    }
  }
  for (;;) {
    double total = 1.0;
    for (int k = 0; k < 3; ++k) {
      double nk = std::floor(ext[k] / h) + 1.0;
      if (nk > 1e9) nk = 1e9;
      g.n[k] = (int)nk;
      total *= nk;
    }
    if (total <= max_cells) break;
    h *= 1.5;
  }
  g.inv_h = 1.0 / h;
  uint64_t n_cells = (uint64_t)g.n[0] * g.n[1] * g.n[2];
  C2B_TRY(ctx->cell_of_pt.ensure(P * 4));
  C2B_TRY(ctx->cell_start.ensure((n_cells + 1) * 4));
  C2B_TRY(ctx->cell_cursor.ensure((n_cells + 1) * 4));
  C2B_TRY(ctx->grid_x.ensure(P * 8));
  C2B_TRY(ctx->grid_y.ensure(P * 8));
  C2B_TRY(ctx->grid_z.ensure(P * 8));
  C2B_TRY(ctx->grid_idx.ensure(P * 4));
  C2B_CUDA(cudaMemsetAsync(ctx->cell_start.p, 0, (n_cells + 1) * 4, st));
  const double *px = ctx->pts.as<double>(), *py = px + P, *pz = py + P;
  k_grid_count<<<blocks_for(P, 256), 256, 0, st>>>(px, py, pz, P, g, ctx->cell_of_pt.as<uint32_t>(),
                                                   ctx->cell_start.as<uint32_t>());
  C2B_KERNEL_CHECK();
  C2B_TRY(exclusive_scan_u32(st, ctx->cell_start.as<uint32_t>(), ctx->cell_start.as<uint32_t>(),
                             n_cells + 1, nullptr, ctx->scan_tmp));
  C2B_CUDA(cudaMemcpyAsync(ctx->cell_cursor.p, ctx->cell_start.p, (n_cells + 1) * 4, cudaMemcpyDeviceToDevice, st));
  k_grid_fill<<<blocks_for(P, 256), 256, 0, st>>>(
      px, py, pz, P, ctx->cell_of_pt.as<uint32_t>(), ctx->cell_cursor.as<uint32_t>(), ctx->grid_x.as<double>(),
      ctx->grid_y.as<double>(), ctx->grid_z.as<double>(), ctx->grid_idx.as<uint32_t>());
  C2B_KERNEL_CHECK();
  x->grid.valid = true;
  x->grid.max_dist = max_dist;
  x->grid.points_version = x->points_version;
  x->grid.desc = g;
  x->grid.n_cells = n_cells;
  return C2B_OK;
}

// the point grid for the resident cameras: the cached one if it still serves them, else a new one
static int ensure_grid(c2b_ctx *ctx, CtxExtra *x, double max_dist) {
  if (!(ctx->C && ctx->P)) return C2B_OK;
  bool valid = x->grid.valid && x->grid.max_dist == max_dist && x->grid.points_version == x->points_version;
  // built for some cameras' reach: still good if the current ones stay inside it (a host-buffer call builds it
  // once for ALL its cameras and locks it for the call's batches)
  if (valid && x->grid.hinted && !x->grid_locked) {
    double clo[3], chi[3], rlo[3], rhi[3];
    C2B_TRY(camera_bbox(ctx, clo, chi));
    reach_box(clo, chi, max_dist, rlo, rhi);
    for (int k = 0; k < 3; ++k) valid = valid && rlo[k] >= x->grid.hint_lo[k] && rhi[k] <= x->grid.hint_hi[k];
  }
  if (!valid) C2B_TRY(build_grid(ctx, x, max_dist));
  return C2B_OK;
}

namespace {

void fill_stats(c2b_ctx *ctx, c2b_obs *stats, uint64_t C, uint64_t n_cand, uint64_t pairs_eval,
                uint64_t nodes, uint64_t tris) {
  if (!stats) return;
  memset(stats, 0, sizeof *stats);
  stats->n_cameras = C;
  stats->n_obs = ctx->out_O;
  stats->n_candidates = n_cand;
  stats->pairs_evaluated = pairs_eval;
  stats->nodes_visited = nodes;
  stats->tris_tested = tris;
  auto ms = [&](int a, int b) {
    float t = 0;
    cudaEventElapsedTime(&t, ctx->ev[a], ctx->ev[b]);
    return t;
  };
  stats->ms_prep = ms(EV_H2D, EV_PREP);
  stats->ms_cull = ms(EV_PREP, EV_CULL);
  stats->ms_sort = ms(EV_CULL, EV_SORT);
  stats->ms_traverse = ms(EV_SORT, EV_TRAVERSE);
  stats->ms_compact = ms(EV_TRAVERSE, EV_COMPACT);
  stats->ms_total = ms(EV_START, EV_D2H);
}

// ---- grid schedule: plan -> fused cull + occlusion -> per-camera sort + write (c2b_fused.cuh) ------
// Stage timers: ms_cull = plan + scan (+ per-camera leaf lists), ms_traverse = the fused kernel,
// ms_compact = visible-count scan + sort/write; ms_sort stays 0 (there is no global sort).
int visibility_grid_fused(c2b_ctx *ctx, CtxExtra *x, const c2b_scene *scene, double max_dist,
                          const c2b_vis_options &opt, int pbits, int cbits, c2b_obs *stats) {
  cudaStream_t st = ctx->stream;
  const uint64_t C = ctx->C, P = ctx->P;
  const int sel = ctx->out_sel;
  C2B_TRY(ensure_grid(ctx, x, max_dist));
  C2B_CUDA(cudaEventRecord(ctx->ev[EV_PREP], st));

  C2B_TRY(ctx->counters.ensure(128));
  C2B_TRY(ctx->out_offsets[sel].ensure((C + 1) * 8));
  C2B_CUDA(cudaMemsetAsync(ctx->counters.p, 0, 128, st));
  uint64_t n_cand = 0, pairs_eval = 0, total_obs = 0, nodes = 0, tris_t = 0;
  if (!(C && P)) {
    C2B_CUDA(cudaMemsetAsync(ctx->out_offsets[sel].p, 0, (C + 1) * 8, st));
    for (int e : {EV_CULL, EV_SORT, EV_TRAVERSE, EV_COMPACT, EV_D2H}) C2B_CUDA(cudaEventRecord(ctx->ev[e], st));
    C2B_CUDA(cudaStreamSynchronize(st));
    ctx->out_C = C;
    ctx->out_O = 0;
    x->have_result = true;
    x->res_candidates = 0;
    fill_stats(ctx, stats, C, 0, 0, 0, 0);
    return C2B_OK;
  }

  // Tickets per camera: with fewer cameras than persistent warps a camera's rows are split over 2 or 4
  // tickets (a ticket repeats the per-camera set-up, so it only pays on warps that would otherwise idle).
  int parts_log2 = 0;
  {
    const uint64_t warps = (uint64_t)ctx->sm_count * FU_RESIDENT_WARPS;
    if (2 * C <= warps) parts_log2 = 1;
    if (4 * C <= warps) parts_log2 = 2;
    if (x->tun.parts_log2 >= 0) parts_log2 = x->tun.parts_log2;  // test hook
  }
  const uint64_t slots = C << parts_log2;
  C2B_TRY(ctx->ev_off.ensure((slots + 1) * 4));
  C2B_TRY(ctx->vis_count.ensure((slots + 1) * 4));
  C2B_TRY(ctx->seg_off.ensure((slots + 1) * 4));
  const double *cxp = ctx->cam_center.as<double>();
  const bool mesh = opt.occlusion == C2B_OCC_MESH && scene && scene->n_nodes > 0;
  FusedArgs fa;
  memset(&fa, 0, sizeof fa);
  fa.cams = ctx->cams.as<double>();
  fa.cen_x = cxp;
  fa.cen_y = cxp + C;
  fa.cen_z = cxp + 2 * C;
  fa.C = C;
  fa.gx = ctx->grid_x.as<double>();
  fa.gy = ctx->grid_y.as<double>();
  fa.gz = ctx->grid_z.as<double>();
  fa.gidx = ctx->grid_idx.as<uint32_t>();
  fa.cell_start = ctx->cell_start.as<uint32_t>();
  fa.g = x->grid.desc;
  fa.max_dist = max_dist;
  fa.t_star = exact_sq_threshold(max_dist);
  fa.scene_absmax = scene_abs_max(scene);
  fa.endpoint_guard_rel = opt.endpoint_guard_rel;
  fa.block_length = opt.block_length;
  fa.block_inset = opt.block_inset;
  fa.parts_log2 = parts_log2;
  C2B_TRY(ctx->plan_rows.ensure(C * FU_ROWS * sizeof(uint2)));
  C2B_TRY(ctx->plan_row_count.ensure(C * 4));
  fa.rows = ctx->plan_rows.as<uint2>();
  fa.row_count = ctx->plan_row_count.as<uint32_t>();
  fa.ev_count = ctx->ev_off.as<uint32_t>();
  fa.vis_count = ctx->vis_count.as<uint32_t>();
  fa.counters = ctx->counters.as<unsigned long long>();

  uint32_t *d_total = ctx->counters.as<uint32_t>() + 12;  // bytes 48..51 of the counter block
  uint32_t *d_max = ctx->counters.as<uint32_t>() + 13;    // bytes 52..55
  unsigned long long h_cnt[16];

  // Per-camera leaf lists (cameras + BVH only) and the plan (cameras + point grid only) are independent and
  // each too small to fill the GPU (one thread or warp per camera, latency-bound): the lists run on a
  // second stream beside the plan and its scan, and the main stream joins before the counters are read.
  cudaStream_t ax = ctx->aux_stream;
  if (mesh) {
    const uint32_t trilist_cap = (uint32_t)x->tun.trilist_cap;
    C2B_TRY(ctx->tri_list.ensure((size_t)C * trilist_cap * 4));
    C2B_TRY(ctx->tri_count.ensure(C * 4));
    float rmax = (float)max_dist;
    if ((double)rmax < max_dist) rmax = std::nextafter(rmax, INFINITY);
    rmax = rmax * (1.0f + 4e-6f);
    fa.nodes = scene->nodes.as<float4>();
    fa.tris = scene->tris.as<float4>();
    fa.n_nodes = (int)scene->n_nodes;
    fa.tri_list = ctx->tri_list.as<uint32_t>();
    fa.tri_count = ctx->tri_count.as<uint32_t>();
    fa.tri_cap = trilist_cap;
    fa.hoist_max = (uint32_t)std::min(x->tun.hoist_max, FU_HOIST);
    fa.packet_bvh = x->tun.packet_bvh;
    const bool tl_warp = x->tun.trilist_warp >= 0 ? x->tun.trilist_warp != 0 : C < 32768;
    C2B_CUDA(cudaEventRecord(ctx->ev_fork, st));  // after the counter reset and the camera upload
    C2B_CUDA(cudaStreamWaitEvent(ax, ctx->ev_fork, 0));
    if (tl_warp)
      k_cam_trilist_warp<<<blocks_for(C, TL_WARPS), TL_WARPS * 32, 0, ax>>>(
          fa.nodes, fa.n_nodes, fa.cen_x, fa.cen_y, fa.cen_z, C, rmax, fa.scene_absmax, trilist_cap,
          ctx->tri_list.as<uint32_t>(), ctx->tri_count.as<uint32_t>(), fa.counters + 0);
    else
      k_cam_trilist<<<blocks_for(C, 128), 128, 0, ax>>>(fa.nodes, fa.n_nodes, fa.cen_x, fa.cen_y, fa.cen_z, C, rmax,
                                                       fa.scene_absmax, trilist_cap, ctx->tri_list.as<uint32_t>(),
                                                       ctx->tri_count.as<uint32_t>(), fa.counters + 0);
    C2B_KERNEL_CHECK();
    C2B_CUDA(cudaEventRecord(ctx->ev_join, ax));
  }
  // plan: row points per camera -> scratch slices
  k_cam_plan<<<blocks_for(C + 1, 128), 128, 0, st>>>(fa);
  C2B_KERNEL_CHECK();
  C2B_TRY(exclusive_scan_u32(st, fa.ev_count, fa.ev_count, slots + 1, nullptr, ctx->scan_tmp));
  if (mesh) C2B_CUDA(cudaStreamWaitEvent(st, ctx->ev_join, 0));
  // Sizes of the scratch / output arrays and the kernel variant depend on the plan's totals.  The first pass
  // waits for them; later passes go by what the previous pass left (grow-only buffers, the remembered variant),
  // launch everything back to back, and check at the one final sync that it all fitted.
  const uint64_t max_pairs = x->tun.max_pairs;  // 32-bit scratch offsets (lower: test hook)
  const bool mt = mesh && opt.predicate == C2B_PRED_MT;
  const bool epi = x->tun.epilogue && parts_log2 == 0 && !opt.count_traversal && !mt && !(mesh && x->tun.fu_occ3);
  const bool optimistic = x->tun.optimistic && x->memo.valid && !epi && !opt.count_traversal && ctx->scratch_idx.cap >= 4 &&
                          ctx->out_idx[sel].cap >= 4 && ctx->out_uv[sel].cap >= 16;
  bool any_overflow = x->memo.any_overflow;
  if (!optimistic) {
    C2B_TRY(read_counters(ctx, h_cnt));
    pairs_eval = h_cnt[1];
    any_overflow = h_cnt[0] != 0;  // cameras whose leaf list exceeded the cap
    if (pairs_eval >= max_pairs)
      return set_error(ERR_TOO_LARGE, "more than 2^32 camera-point pairs inside max_dist in one call; shard the cameras");
    C2B_TRY(ctx->scratch_idx.ensure(std::max<uint64_t>(pairs_eval, 1) * 4));
  }
  fa.scratch_idx = ctx->scratch_idx.as<uint32_t>();
  fa.scratch_cap = ctx->scratch_idx.cap / 4;
  C2B_CUDA(cudaEventRecord(ctx->ev[EV_CULL], st));
  C2B_CUDA(cudaEventRecord(ctx->ev[EV_SORT], st));

  // fused cull + occlusion (+ the in-kernel sort / CSR write: see epilogue_sort_write)
  C2B_CUDA(cudaMemsetAsync(ctx->vis_count.p, 0, (slots + 1) * 4, st));
  if (epi) {
    // the visible count is at most the planned row points: output arrays of that size can never overflow
    C2B_TRY(ctx->out_idx[sel].ensure(std::max<uint64_t>(pairs_eval, 1) * 4));
    C2B_TRY(ctx->out_uv[sel].ensure(std::max<uint64_t>(pairs_eval, 1) * 16));
    const uint64_t c_pad = (C + 127) & ~127ull;  // the scanner reads and writes whole 128-camera steps
    C2B_TRY(ctx->epi_status.ensure(c_pad * 4));
    C2B_TRY(ctx->epi_prefix.ensure(c_pad * 8));
    C2B_CUDA(cudaMemsetAsync(ctx->epi_status.p, 0, c_pad * 4, st));
    C2B_CUDA(cudaMemsetAsync(ctx->epi_prefix.p, 0, c_pad * 8, st));
    fa.status = ctx->epi_status.as<uint32_t>();
    fa.prefix = ctx->epi_prefix.as<unsigned long long>();
    fa.p_aos = ctx->pts_aos.as<double>();
    fa.out_offsets = ctx->out_offsets[sel].as<uint64_t>();
    fa.out_idx = ctx->out_idx[sel].as<uint32_t>();
    fa.out_uv = ctx->out_uv[sel].as<double2>();
    fa.out_cap = std::min<uint64_t>(ctx->out_idx[sel].cap / 4, ctx->out_uv[sel].cap / 16);
    fa.key_bits = std::max(pbits, 1);
  }
  {
    // persistent warps draw cameras from a ticket; more CTAs than can be resident is harmless
    const unsigned nb = (unsigned)std::min<uint64_t>(blocks_for(slots, FU_WARPS), (uint64_t)ctx->sm_count * 8), nt = FU_WARPS * 32;
    const bool cnt = opt.count_traversal != 0;
    // The template's third argument counts CTAs of EIGHT warps per SM (fu_min_ctas translates for four-warp CTAs).
    // Eight-warp CTAs, r01m kernels at cfg4: 4 CTAs/SM at 64 registers 2.68 ms; 3 at 80 (C2B_FU_OCC3) 2.69 ms; 5 at 48
    // 2.74 ms.  Four-warp CTAs, r02y kernels: 7 CTAs/SM at 72 registers 2.31 ms against 2.38 ms for 4 x 8 warps at 64.
    const bool occ4 = !x->tun.fu_occ3;
    if (mt) {
      // candidates only; k_filter_candidates_mt decides occlusion below
      k_visibility_fused<FU_OCC_NONE, false, 3, false><<<nb, nt, 0, st>>>(fa);
    } else if (mesh) {
      if (cnt)
        k_visibility_fused<FU_OCC_MESH, true, 3, true><<<nb, nt, 0, st>>>(fa);
      else if (any_overflow && epi)
        k_visibility_fused<FU_OCC_MESH, false, 4, true, true><<<nb + 1, nt, 0, st>>>(fa);
      else if (any_overflow)
        k_visibility_fused<FU_OCC_MESH, false, 4, true><<<nb, nt, 0, st>>>(fa);
      else if (epi)
        k_visibility_fused<FU_OCC_MESH, false, 4, false, true><<<nb + 1, nt, 0, st>>>(fa);
      else if (occ4)
        k_visibility_fused<FU_OCC_MESH, false, 4, false><<<nb, nt, 0, st>>>(fa);
      else
        k_visibility_fused<FU_OCC_MESH, false, 3, false><<<nb, nt, 0, st>>>(fa);
    } else if (opt.occlusion == C2B_OCC_ANALYTIC) {
      if (epi)
        k_visibility_fused<FU_OCC_ANALYTIC, false, 2, false, true><<<nb + 1, nt, 0, st>>>(fa);
      else
        k_visibility_fused<FU_OCC_ANALYTIC, false, 2, false><<<nb, nt, 0, st>>>(fa);
    } else {
      if (epi)
        k_visibility_fused<FU_OCC_NONE, false, 3, false, true><<<nb + 1, nt, 0, st>>>(fa);
      else
        k_visibility_fused<FU_OCC_NONE, false, 3, false><<<nb, nt, 0, st>>>(fa);
    }
    C2B_KERNEL_CHECK();
    if (mt) {
      FilterArgs f{fa.nodes, fa.tris, fa.n_nodes, fa.scene_absmax, fa.cen_x, fa.cen_y, fa.cen_z, ctx->pts_aos.as<double>(),
                   slots, parts_log2, opt.endpoint_guard_rel, fa.ev_count, fa.scratch_idx, fa.vis_count, fa.counters};
      k_filter_candidates_mt<<<blocks_for(slots, 8), 256, 0, st>>>(f);
      C2B_KERNEL_CHECK();
    }
  }
  C2B_CUDA(cudaEventRecord(ctx->ev[EV_TRAVERSE], st));

  bool epi_done = false;
  if (epi) {
    // the kernel has written the CSR; the counters say whether every camera took part
    C2B_TRY(read_counters(ctx, h_cnt));
    if (h_cnt[5]) return set_error(C2B_ERR_CUDA, "internal: a camera's scratch slice overflowed");
    if (h_cnt[11]) return set_error(C2B_ERR_CUDA, "internal: the fused kernel's epilogue gave up waiting for a camera's count");
    n_cand = h_cnt[4];
    if (h_cnt[9] <= SW_WARP_MAX && h_cnt[10] == 0) {
      total_obs = h_cnt[8];
      epi_done = true;
    }
    // else: a camera sees more points than the warp sort holds — the two-kernel path below redoes the tail
  }
  if (!epi_done) {
  // visible counts -> CSR offsets
  C2B_TRY(exclusive_scan_u32(st, fa.vis_count, ctx->seg_off.as<uint32_t>(), slots + 1, d_total, ctx->scan_tmp));
  k_max_cam<<<(unsigned)std::min<uint64_t>(blocks_for(C, 256), 1024), 256, 0, st>>>(ctx->seg_off.as<uint32_t>(), C, parts_log2, d_max);
  C2B_KERNEL_CHECK();
  uint32_t total32 = 0, max32 = 0;
  auto read_totals = [&]() -> int {
    C2B_TRY(read_counters(ctx, h_cnt));
    if (h_cnt[5]) return set_error(C2B_ERR_CUDA, "internal: a camera's scratch slice overflowed");
    memcpy(&total32, reinterpret_cast<const char *>(h_cnt) + 48, 4);
    memcpy(&max32, reinterpret_cast<const char *>(h_cnt) + 52, 4);
    total_obs = total32;
    n_cand = h_cnt[4];
    nodes = h_cnt[2];
    tris_t = h_cnt[3];
    return C2B_OK;
  };
  auto sort_write_args = [&]() {
    return SortWriteArgs{fa.ev_count, ctx->seg_off.as<uint32_t>(), C, fa.scratch_idx, fa.cams, ctx->pts_aos.as<double>(),
                         ctx->out_offsets[sel].as<uint64_t>(), ctx->out_idx[sel].as<uint32_t>(), ctx->out_uv[sel].as<double2>(),
                         std::max(pbits, 1), parts_log2,
                         std::min<uint64_t>(ctx->out_idx[sel].cap / 4, ctx->out_uv[sel].cap / 16), fa.counters + 12};
  };
  const unsigned swb = (unsigned)blocks_for(C, SW_WARPS), swt = SW_WARPS * 32;
  if (optimistic) {
    // sort + write straight away into the arrays as they are; the one sync of the pass comes after it
    const SortWriteArgs sw = sort_write_args();
    if (parts_log2 > 0)
      k_sort_write<6, true><<<swb, swt, 0, st>>>(sw);
    else
      k_sort_write<8, false><<<swb, swt, 0, st>>>(sw);
    C2B_KERNEL_CHECK();
    if (x->memo.big) {
      k_sort_write_block<<<(unsigned)C, 256, 0, st>>>(sw);
      C2B_KERNEL_CHECK();
    }
    C2B_TRY(read_totals());
    pairs_eval = h_cnt[1];
    const bool fits = h_cnt[12] == 0 && pairs_eval <= fa.scratch_cap && pairs_eval < max_pairs &&
                      (max32 <= SW_WARP_MAX || (x->memo.big && max32 <= SW_BLOCK_MAX));
    if (!fits) {
      // this pass is larger than what the previous one left behind: size everything exactly and repeat it
      x->memo.valid = false;
      if (pairs_eval >= max_pairs)
        return set_error(ERR_TOO_LARGE, "more than 2^32 camera-point pairs inside max_dist in one call; shard the cameras");
      return visibility_grid_fused(ctx, x, scene, max_dist, opt, pbits, cbits, stats);
    }
    x->memo.any_overflow = h_cnt[0] != 0;
    x->memo.big = max32 > SW_WARP_MAX;
  } else {
  C2B_TRY(read_totals());
  x->memo.valid = true;
  x->memo.any_overflow = any_overflow;
  x->memo.big = max32 > SW_WARP_MAX;
  if (total_obs) {
    C2B_TRY(ctx->out_idx[sel].ensure(total_obs * 4));
    C2B_TRY(ctx->out_uv[sel].ensure(total_obs * 16));
    const SortWriteArgs sw = sort_write_args();
    if (max32 <= SW_BLOCK_MAX) {
      // registers capped for 8 CTAs/SM (measured at cfg4: 0.81 ms; 6 CTAs/SM 0.91 ms, 4 CTAs/SM 1.12 ms)
      if (parts_log2 > 0)
        k_sort_write<6, true><<<swb, swt, 0, st>>>(sw);
      else
        k_sort_write<8, false><<<swb, swt, 0, st>>>(sw);
      C2B_KERNEL_CHECK();
      if (max32 > SW_WARP_MAX) {
        k_sort_write_block<<<(unsigned)C, 256, 0, st>>>(sw);
        C2B_KERNEL_CHECK();
      }
    } else {
      // a camera sees more points than the shared-memory sort holds: radix-sort all visible keys
      C2B_TRY(ctx->sort_keys[0].ensure(total_obs * 8));
      C2B_TRY(ctx->sort_keys[1].ensure(total_obs * 8));
      C2B_TRY(ctx->sort_vals[0].ensure(total_obs * 4));
      C2B_TRY(ctx->sort_vals[1].ensure(total_obs * 4));
      k_expand_keys<<<blocks_for(C, 8), 256, 0, st>>>(sw, pbits, ctx->sort_keys[0].as<uint64_t>());
      C2B_KERNEL_CHECK();
      uint64_t *keys[2] = {ctx->sort_keys[0].as<uint64_t>(), ctx->sort_keys[1].as<uint64_t>()};
      uint32_t *vals[2] = {ctx->sort_vals[0].as<uint32_t>(), ctx->sort_vals[1].as<uint32_t>()};
      int res = 0;
      C2B_TRY(radix_sort_pairs(st, keys, vals, total_obs, pbits + cbits, ctx->sort_hist, ctx->scan_tmp, &res));
      k_write_sorted<<<blocks_for(total_obs, 256), 256, 0, st>>>(keys[res], total_obs, pbits, fa.cams,
                                                                ctx->pts_aos.as<double>(), ctx->out_idx[sel].as<uint32_t>(),
                                                                ctx->out_uv[sel].as<double2>());
      C2B_KERNEL_CHECK();
      k_widen_cam_offsets<<<blocks_for(C + 1, 256), 256, 0, st>>>(ctx->seg_off.as<uint32_t>(), C + 1, parts_log2,
                                                                 ctx->out_offsets[sel].as<uint64_t>());
      C2B_KERNEL_CHECK();
    }
  } else {
    C2B_CUDA(cudaMemsetAsync(ctx->out_offsets[sel].p, 0, (C + 1) * 8, st));
  }
  }  // !optimistic
  }  // !epi_done
  C2B_CUDA(cudaEventRecord(ctx->ev[EV_COMPACT], st));
  C2B_CUDA(cudaEventRecord(ctx->ev[EV_D2H], st));
  C2B_CUDA(cudaStreamSynchronize(st));
  ctx->out_C = C;
  ctx->out_O = total_obs;
  x->have_result = true;
  x->res_candidates = n_cand;
  fill_stats(ctx, stats, C, n_cand, pairs_eval, nodes, tris_t);
  return C2B_OK;
}

}  // namespace

static int visibility_resident_impl(c2b_ctx *ctx, const c2b_scene *scene, double max_dist,
                                    const c2b_vis_options *opt_in, c2b_obs *stats);

int c2b_visibility_graph_resident(c2b_ctx *ctx, const c2b_scene *scene, double max_dist,
                                  const c2b_vis_options *opt_in, c2b_obs *stats) {
  const int rc = visibility_resident_impl(ctx, scene, max_dist, opt_in, stats);
  return rc == ERR_TOO_LARGE ? C2B_ERR_INVALID : rc;
}

static int visibility_resident_impl(c2b_ctx *ctx, const c2b_scene *scene, double max_dist,
                                    const c2b_vis_options *opt_in, c2b_obs *stats) {
  if (!ctx) return set_error(C2B_ERR_INVALID, "c2b_visibility_graph_resident: null ctx");
  CtxExtra *x = extra_of(ctx);
  if (!x->have_points || !x->have_cameras)
    return set_error(C2B_ERR_INVALID, "upload cameras and points first");
  c2b_vis_options opt;
  if (opt_in)
    opt = *opt_in;
  else
    c2b_vis_options_default(&opt);
  if (opt.occlusion == C2B_OCC_MESH && !scene)
    return set_error(C2B_ERR_INVALID, "C2B_OCC_MESH needs a scene");
  if (opt.cull_mode != C2B_CULL_GRID && opt.cull_mode != C2B_CULL_EXHAUSTIVE)
    return set_error(C2B_ERR_INVALID, "unknown cull_mode %d", opt.cull_mode);
  if (opt.occlusion < C2B_OCC_MESH || opt.occlusion > C2B_OCC_ANALYTIC)
    return set_error(C2B_ERR_INVALID, "unknown occlusion %d", opt.occlusion);
  if (opt.predicate != C2B_PRED_WATERTIGHT && opt.predicate != C2B_PRED_MT)
    return set_error(C2B_ERR_INVALID, "unknown predicate %d", opt.predicate);
  C2B_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const uint64_t C = ctx->C, P = ctx->P;
  x->have_result = false;

  C2B_CUDA(cudaEventRecord(ctx->ev[EV_START], st));
  C2B_CUDA(cudaEventRecord(ctx->ev[EV_H2D], st));

  const int pbits = std::max(1, bit_length(P ? P - 1 : 0));
  const int cbits = std::max(1, bit_length(C ? C - 1 : 0));
  if (pbits + cbits > 63) return set_error(C2B_ERR_INVALID, "camera x point index space exceeds 63 bits");
  if (opt.cull_mode == C2B_CULL_GRID) return visibility_grid_fused(ctx, x, scene, max_dist, opt, pbits, cbits, stats);

  // ---- exhaustive schedule: every pair -> pool -> radix sort -> ordered traversal -> stream compaction --
  C2B_TRY(ctx->counters.ensure(128));
  C2B_TRY(ctx->cam_count.ensure((C + 1) * 4));
  C2B_TRY(ctx->out_offsets[ctx->out_sel].ensure((C + 1) * 8));
  C2B_CUDA(cudaEventRecord(ctx->ev[EV_PREP], st));

  uint64_t n_cand = 0, pairs_eval = 0, pool_n = 0;
  if (C && P) {
    CullArgs a;
    a.cams = ctx->cams.as<double>();
    a.cen_x = ctx->cam_center.as<double>();
    a.cen_y = a.cen_x + C;
    a.cen_z = a.cen_y + C;
    a.C = C;
    a.P = P;
    a.t_star = exact_sq_threshold(max_dist);
    a.t_cons = a.t_star * (1.0 + 4e-15);
    if (a.t_star > 0.0 && a.t_cons == a.t_star) a.t_cons = std::nextafter(a.t_star, INFINITY);
    a.pbits = pbits;
    a.cam_count = ctx->cam_count.as<uint32_t>();
    a.counters = ctx->counters.as<unsigned long long>();
    a.px = ctx->pts.as<double>();
    a.py = a.px + P;
    a.pz = a.py + P;
    if (ctx->pool_capacity == 0) {
      long double pairs = (long double)C * (long double)P;
      uint64_t guess = std::max<uint64_t>(1u << 20, 96 * C);
      if ((long double)guess > pairs) guess = (uint64_t)pairs;
      ctx->pool_capacity = std::max<uint64_t>(guess, 1024);
    }
    for (int attempt = 0;; ++attempt) {
      // the pool's key array doubles as buffer 0 of the radix sort
      C2B_TRY(ctx->sort_keys[0].ensure(ctx->pool_capacity * 8));
      C2B_TRY(ctx->pool_uv.ensure(ctx->pool_capacity * 16));
      a.pool_key = ctx->sort_keys[0].as<uint64_t>();
      a.pool_uv = ctx->pool_uv.as<double2>();
      a.pool_capacity = ctx->pool_capacity;
      C2B_CUDA(cudaMemsetAsync(ctx->counters.p, 0, 64, st));
      C2B_CUDA(cudaMemsetAsync(ctx->cam_count.p, 0, (C + 1) * 4, st));
      dim3 grid(blocks_for(P, CB_THREADS * CB_PPT), blocks_for(C, CB_TC));
      if (grid.y > 65535u) return set_error(C2B_ERR_INVALID, "too many cameras for one exhaustive launch");
      k_cull_exhaustive<<<grid, CB_THREADS, 0, st>>>(a);
      C2B_KERNEL_CHECK();
      unsigned long long h_cnt[16];
      C2B_TRY(read_counters(ctx, h_cnt));
      pool_n = h_cnt[0];
      n_cand = h_cnt[0];
      pairs_eval = C * P;
      if (pool_n <= ctx->pool_capacity) break;
      if (attempt >= 2) return set_error(C2B_ERR_CUDA, "candidate pool overflow persisted");
      ctx->pool_capacity = pool_n + pool_n / 16 + 1024;  // exact count is known now: re-run once
    }
    if (pool_n >= 0xffffffffull)
      return set_error(ERR_TOO_LARGE, "more than 2^32 candidates in one call; shard the cameras");
  } else {
    C2B_CUDA(cudaMemsetAsync(ctx->counters.p, 0, 64, st));
    C2B_CUDA(cudaMemsetAsync(ctx->cam_count.p, 0, (C + 1) * 4, st));
  }
  C2B_CUDA(cudaEventRecord(ctx->ev[EV_CULL], st));

  const double *cxp = ctx->cam_center.as<double>();
  const double *pxp = ctx->pts.as<double>();
  const float scene_absmax = scene_abs_max(scene);

  // occlusion of `n` sorted candidate keys -> one ballot word per 32 keys
  auto run_occlusion = [&](const uint64_t *keys_in, uint64_t n, uint64_t n_words) -> int {
    if (!n) return C2B_OK;
    if (opt.occlusion == C2B_OCC_MESH && scene->n_nodes > 0) {
      TraverseArgs t;
      t.nodes = scene->nodes.as<float4>();
      t.tris = scene->tris.as<float4>();
      t.n_nodes = (int)scene->n_nodes;
      t.keys = keys_in;
      t.n_cand = n;
      t.scene_absmax = scene_absmax;
      t.pbits = pbits;
      t.cen_x = cxp;
      t.cen_y = cxp + C;
      t.cen_z = cxp + 2 * C;
      t.p_aos = ctx->pts_aos.as<double>();
      t.endpoint_guard_rel = opt.endpoint_guard_rel;
      t.vis_words = ctx->vis_words.as<uint32_t>();
      t.counters = ctx->counters.as<unsigned long long>();
      if (opt.predicate == C2B_PRED_MT)
        k_traverse<false, true><<<blocks_for(n_words, 8), 256, 0, st>>>(t);
      else if (opt.count_traversal)
        k_traverse<true><<<blocks_for(n_words, 8), 256, 0, st>>>(t);
      else
        k_traverse<false><<<blocks_for(n_words, 8), 256, 0, st>>>(t);
    } else if (opt.occlusion == C2B_OCC_ANALYTIC) {
      k_analytic_occlusion<<<blocks_for(n_words * 32, 256), 256, 0, st>>>(
          keys_in, n, pbits, cxp, cxp + C, cxp + 2 * C, pxp, pxp + P, pxp + 2 * P, opt.block_length,
          opt.block_inset, ctx->vis_words.as<uint32_t>());
    } else {
      k_words_all_visible<<<blocks_for(n_words * 32, 256), 256, 0, st>>>(keys_in, ctx->vis_words.as<uint32_t>(), n);
    }
    C2B_KERNEL_CHECK();
    return C2B_OK;
  };

  uint32_t *d_total = ctx->counters.as<uint32_t>() + 12;  // bytes 48..51 of the counter block
  uint64_t total_obs = 0;
  unsigned long long h_fin[16] = {0};
  {
    C2B_TRY(exclusive_scan_u32(st, ctx->cam_count.as<uint32_t>(), ctx->cam_count.as<uint32_t>(), C + 1,
                               nullptr, ctx->scan_tmp));
    int res = 0;
    uint64_t *keys[2] = {nullptr, nullptr};
    uint32_t *vals[2] = {nullptr, nullptr};
    if (n_cand) {
      C2B_TRY(ctx->sort_keys[1].ensure(n_cand * 8));
      C2B_TRY(ctx->sort_vals[0].ensure(n_cand * 4));
      C2B_TRY(ctx->sort_vals[1].ensure(n_cand * 4));
      keys[0] = ctx->sort_keys[0].as<uint64_t>();
      keys[1] = ctx->sort_keys[1].as<uint64_t>();
      vals[0] = ctx->sort_vals[0].as<uint32_t>();
      vals[1] = ctx->sort_vals[1].as<uint32_t>();
      k_iota_u32<<<blocks_for(n_cand, 256), 256, 0, st>>>(vals[0], n_cand);
      C2B_KERNEL_CHECK();
      C2B_TRY(radix_sort_pairs(st, keys, vals, n_cand, pbits + cbits, ctx->sort_hist, ctx->scan_tmp, &res));
    }
    C2B_CUDA(cudaEventRecord(ctx->ev[EV_SORT], st));

    const uint64_t n_words = (n_cand + 31) / 32;
    C2B_TRY(ctx->vis_words.ensure((n_words + 1) * 4));
    C2B_TRY(ctx->word_prefix.ensure((n_words + 1) * 4));
    C2B_TRY(run_occlusion(keys[res], n_cand, n_words));
    C2B_CUDA(cudaEventRecord(ctx->ev[EV_TRAVERSE], st));

    C2B_CUDA(cudaMemsetAsync(d_total, 0, 8, st));
    if (n_cand) {
      C2B_TRY(ctx->out_idx[ctx->out_sel].ensure(n_cand * 4));
      C2B_TRY(ctx->out_uv[ctx->out_sel].ensure(n_cand * 16));
      k_word_popc<<<blocks_for(n_words, 256), 256, 0, st>>>(ctx->vis_words.as<uint32_t>(), n_words,
                                                            ctx->word_prefix.as<uint32_t>());
      C2B_KERNEL_CHECK();
      C2B_TRY(exclusive_scan_u32(st, ctx->word_prefix.as<uint32_t>(), ctx->word_prefix.as<uint32_t>(),
                                 n_words, d_total, ctx->scan_tmp));
      k_compact_write<<<blocks_for(n_cand, 256), 256, 0, st>>>(
          ctx->vis_words.as<uint32_t>(), ctx->word_prefix.as<uint32_t>(), keys[res], vals[res],
          ctx->pool_uv.as<double2>(), n_cand, pbits, ctx->out_idx[ctx->out_sel].as<uint32_t>(),
          ctx->out_uv[ctx->out_sel].as<double2>());
      C2B_KERNEL_CHECK();
    }
    k_csr_offsets<<<blocks_for(C + 1, 256), 256, 0, st>>>(
        ctx->cam_count.as<uint32_t>(), C, ctx->vis_words.as<uint32_t>(), ctx->word_prefix.as<uint32_t>(),
        n_cand, d_total, ctx->out_offsets[ctx->out_sel].as<uint64_t>());
    C2B_KERNEL_CHECK();
    C2B_CUDA(cudaEventRecord(ctx->ev[EV_COMPACT], st));
    C2B_CUDA(cudaEventRecord(ctx->ev[EV_D2H], st));
    C2B_TRY(read_counters(ctx, h_fin));
    uint32_t total32;
    memcpy(&total32, reinterpret_cast<const char *>(h_fin) + 48, 4);
    total_obs = total32;
  }
  ctx->out_C = C;
  ctx->out_O = total_obs;
  x->have_result = true;
  x->res_candidates = n_cand;
  fill_stats(ctx, stats, C, n_cand, pairs_eval, h_fin[2], h_fin[3]);
  return C2B_OK;
}

int c2b_download_obs(c2b_ctx *ctx, c2b_obs *out) {
  if (!ctx || !out) return set_error(C2B_ERR_INVALID, "c2b_download_obs: null argument");
  CtxExtra *x = extra_of(ctx);
  if (!x->have_result) return set_error(C2B_ERR_INVALID, "no resident result to download");
  C2B_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const uint64_t C = ctx->out_C, O = ctx->out_O;
  C2B_TRY(ctx->h_offsets.ensure((C + 1) * 8));
  C2B_TRY(ctx->h_idx.ensure(std::max<uint64_t>(O, 1) * 4));
  C2B_TRY(ctx->h_uv.ensure(std::max<uint64_t>(O, 1) * 16));
  cudaEvent_t e0 = ctx->ev[EV_COMPACT], e1 = ctx->ev[EV_D2H];
  C2B_CUDA(cudaEventRecord(e0, st));
  C2B_CUDA(cudaMemcpyAsync(ctx->h_offsets.p, ctx->out_offsets[ctx->out_sel].p, (C + 1) * 8, cudaMemcpyDeviceToHost, st));
  if (O) {
    C2B_CUDA(cudaMemcpyAsync(ctx->h_idx.p, ctx->out_idx[ctx->out_sel].p, O * 4, cudaMemcpyDeviceToHost, st));
    C2B_CUDA(cudaMemcpyAsync(ctx->h_uv.p, ctx->out_uv[ctx->out_sel].p, O * 16, cudaMemcpyDeviceToHost, st));
  }
  C2B_CUDA(cudaEventRecord(e1, st));
  C2B_CUDA(cudaStreamSynchronize(st));
  out->n_cameras = C;
  out->n_obs = O;
  out->offsets = ctx->h_offsets.as<uint64_t>();
  out->point_idx = ctx->h_idx.as<uint32_t>();
  out->uv = ctx->h_uv.as<double>();
  out->d2h_bytes = (C + 1) * 8 + O * 20;
  float t = 0;
  cudaEventElapsedTime(&t, e0, e1);
  out->ms_d2h = t;
  return C2B_OK;
}

namespace {
__global__ void k_add_u64(uint64_t *v, uint64_t n, uint64_t add) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] += add;
}
}  // namespace

int c2b_download_obs_into(c2b_ctx *ctx, uint64_t obs_base, uint64_t *offsets_dst, uint32_t *idx_dst, double *uv_dst,
                          int with_end, float *ms_d2h) {
  if (!ctx || !offsets_dst) return set_error(C2B_ERR_INVALID, "c2b_download_obs_into: null argument");
  CtxExtra *x = extra_of(ctx);
  if (!x->have_result) return set_error(C2B_ERR_INVALID, "no resident result to download");
  const uint64_t C = ctx->out_C, O = ctx->out_O;
  if (O && (!idx_dst || !uv_dst)) return set_error(C2B_ERR_INVALID, "c2b_download_obs_into: null result array");
  C2B_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const int sel = ctx->out_sel;
  if (obs_base) {
    k_add_u64<<<blocks_for(C + 1, 256), 256, 0, st>>>(ctx->out_offsets[sel].as<uint64_t>(), C + 1, obs_base);
    C2B_KERNEL_CHECK();
  }
  cudaEvent_t e0 = ctx->ev[EV_COMPACT], e1 = ctx->ev[EV_D2H];
  C2B_CUDA(cudaEventRecord(e0, st));
  C2B_CUDA(cudaMemcpyAsync(offsets_dst, ctx->out_offsets[sel].p, (C + (with_end ? 1 : 0)) * 8, cudaMemcpyDeviceToHost, st));
  if (O) {
    C2B_CUDA(cudaMemcpyAsync(idx_dst, ctx->out_idx[sel].p, O * 4, cudaMemcpyDeviceToHost, st));
    C2B_CUDA(cudaMemcpyAsync(uv_dst, ctx->out_uv[sel].p, O * 16, cudaMemcpyDeviceToHost, st));
  }
  C2B_CUDA(cudaEventRecord(e1, st));
  C2B_CUDA(cudaStreamSynchronize(st));
  if (obs_base) {  // leave the resident CSR as it was (a second download must not rebase twice)
    k_add_u64<<<blocks_for(C + 1, 256), 256, 0, st>>>(ctx->out_offsets[sel].as<uint64_t>(), C + 1, (uint64_t)0 - obs_base);
    C2B_KERNEL_CHECK();
  }
  if (ms_d2h) {
    float t = 0;
    cudaEventElapsedTime(&t, e0, e1);
    *ms_d2h = t;
  }
  return C2B_OK;
}

namespace {

// grow a pinned result buffer while copies into it may be in flight: drain, allocate, carry over
int grow_pinned(c2b_ctx *ctx, PinBuf &b, size_t need, size_t used) {
  if (need <= b.cap) return C2B_OK;
  C2B_CUDA(cudaStreamSynchronize(ctx->copy_stream));
  PinBuf nb;
  nb.node = b.node;
  C2B_TRY(nb.ensure(need + need / 16));
  if (used) memcpy(nb.p, b.p, used);
  b.release();
  b = nb;
  return C2B_OK;
}
}  // namespace

int c2b_visibility_graph(c2b_ctx *ctx, const c2b_scene *scene, const double *cams, uint64_t C,
                         const double *pts, uint64_t P, double max_dist, const c2b_vis_options *opt,
                         c2b_obs *out) {
  if (!ctx || !out) return set_error(C2B_ERR_INVALID, "c2b_visibility_graph: null argument");
  if (C && !cams) return set_error(C2B_ERR_INVALID, "c2b_visibility_graph: null camera array");
  C2B_CUDA(cudaSetDevice(ctx->device));
  CtxExtra *x = extra_of(ctx);
  // pts == NULL: use the points already resident on the device (c2b_upload_points[_device])
  const bool resident_pts = P && !pts;
  if (resident_pts && !(x->have_points && ctx->P == P))
    return set_error(C2B_ERR_INVALID, "c2b_visibility_graph: pts is null and no %llu resident points were uploaded",
                     (unsigned long long)P);
  cudaStream_t st = ctx->stream, cs = ctx->copy_stream;
  // Camera batches: the CSR slab of batch b travels to the host while batch b+1 is computed.  A batch should
  // give every resident warp of the fused kernel more than one camera (a short batch leaves warps idle at its
  // tail and repeats the per-batch plan / scan / sync), and:
  //   * when the result transfer is the longest leg (20 B per observation over PCIe: cfg4) it should start
  //     early — a small first batch, then growing ones;
  //   * when the kernels are (fine meshes: cfg5) the batches should be few and equal — only the last slab's
  //     transfer is exposed.
  // Which case applies is taken from the previous call on this ctx; the first call assumes the transfer.
  std::vector<uint64_t> bounds;  // batch b = cameras [bounds[b], bounds[b+1])
  bounds.push_back(0);
  const uint64_t W = (uint64_t)ctx->sm_count * FU_RESIDENT_WARPS;  // resident warps of the fused kernel
  if (x->tun.batches > 0) {
    const uint64_t nb = std::max<uint64_t>(1, std::min<uint64_t>(C ? C : 1, (uint64_t)x->tun.batches));
    for (uint64_t b = 1; b <= nb; ++b) bounds.push_back((b * C) / nb);
  } else if (C < 3 * W) {
    bounds.push_back(C);
  } else if (x->hist.valid && x->hist.ms_compute_per_cam > x->hist.d2h_bytes_per_cam / 52e6) {
    const uint64_t nb = std::min<uint64_t>(std::max<uint64_t>(C / (3 * W), 2), 4);
    for (uint64_t b = 1; b <= nb; ++b) bounds.push_back((b * C) / nb);
  } else {
    uint64_t done = 0, size = std::max<uint64_t>(3 * W / 2, C / 32);
    const uint64_t cap = std::max<uint64_t>(3 * W, C / 6);
    while (C - done > size + size / 2) {
      done += size;
      bounds.push_back(done);
      size = std::min(2 * size, cap);
    }
    bounds.push_back(C);
  }
  const uint64_t n_batches = bounds.size() - 1;

  cudaEvent_t u0, u1, d0, d1;
  C2B_CUDA(cudaEventCreate(&u0));
  C2B_CUDA(cudaEventCreate(&u1));
  C2B_CUDA(cudaEventCreate(&d0));
  C2B_CUDA(cudaEventCreate(&d1));
  c2b_obs acc;
  memset(&acc, 0, sizeof acc);
  int rc = C2B_OK;
  uint64_t obs_base = 0;
  float ms_upload = 0, ms_copy = 0;
  // First host-buffer call on this ctx: the CSR goes to unpinned memory through the Drainer (see PageBuf).
  const bool staged = x->tun.cold_staged && x->tun.stage_threads > 0 && !x->hist.valid && ctx->h_uv.cap < (64u << 20);
  Drainer drain;
  std::chrono::steady_clock::time_point t_first_copy;
  bool copy_started = false;
  if (!staged) {
    x->pg_idx.release();
    x->pg_uv.release();
  }
  auto run = [&]() -> int {
    if (staged) {
      const int nthreads = std::max(2, x->tun.stage_threads);
      x->pin_out.node = ctx->numa_node;
      C2B_TRY(x->pin_out.ensure((size_t)nthreads * 2 * Drainer::CH));
      C2B_TRY(drain.start(ctx->device, nthreads, x->pin_out.as<char>()));
    }
    C2B_CUDA(cudaEventRecord(u0, st));
    if (!resident_pts) C2B_TRY(c2b_upload_points(ctx, pts, P));
    C2B_CUDA(cudaEventRecord(u1, st));
    // the point grid once, for the reach of ALL of this call's cameras (each batch's reach lies inside it)
    const bool grid_mode = !opt || opt->cull_mode == C2B_CULL_GRID;
    // all cameras cross PCIe once; the batches are device-side slices of that copy
    if (C) {
      C2B_TRY(ctx->cams_all.ensure(C * 120));
      C2B_TRY(copy_in(ctx, ctx->cams_all.p, cams, C * 120, cudaMemcpyHostToDevice));
    }
    if (grid_mode && n_batches > 1 && C && P) {
      C2B_TRY(select_cameras(ctx, 0, C));
      C2B_TRY(ensure_grid(ctx, x, max_dist));
      x->grid_locked = true;
    }
    C2B_TRY(ctx->h_offsets.ensure((C + 1) * 8));
    // camera ranges still to do, in order; a range that turns out too large for 32-bit offsets is halved
    std::vector<std::pair<uint64_t, uint64_t>> todo;
    for (uint64_t b = n_batches; b-- > 0;) todo.push_back({bounds[b], bounds[b + 1]});
    for (uint64_t b = 0; !todo.empty(); ++b) {
      const uint64_t c0 = todo.back().first, c1 = todo.back().second, nc = c1 - c0;
      todo.pop_back();
      const int sel = (int)(b & 1);
      ctx->out_sel = sel;
      if (b >= 2) {  // set `sel` is free again once batch b - 2 has left the device
        if (staged)
          drain.wait_dma((int)b - 2);
        else
          C2B_CUDA(cudaStreamWaitEvent(st, ctx->ev_copied[sel], 0));
      }
      C2B_TRY(select_cameras(ctx, c0, nc));
      c2b_obs s1;
      const int rcb = visibility_resident_impl(ctx, scene, max_dist, opt, &s1);
      if (rcb == ERR_TOO_LARGE && nc > 1) {
        todo.push_back({c0 + nc / 2, c1});
        todo.push_back({c0, c0 + nc / 2});
        --b;  // the buffer set was not used
        continue;
      }
      if (rcb != C2B_OK) return rcb == ERR_TOO_LARGE ? C2B_ERR_INVALID : rcb;
      const uint64_t O = ctx->out_O;
      if (obs_base)
        k_add_u64<<<blocks_for(nc + 1, 256), 256, 0, st>>>(ctx->out_offsets[sel].as<uint64_t>(), nc + 1, obs_base);
      ++launch_counter();
      C2B_CUDA(cudaEventRecord(ctx->ev_ready[sel], st));
      // result arrays: sized from the first batch's density, grown if the guess was short
      size_t need = obs_base + O;
      if (!todo.empty()) need = std::max<size_t>(need, (size_t)((double)(obs_base + O) * (double)C / (double)c1 * 1.05) + 1024);
      C2B_CUDA(cudaStreamWaitEvent(cs, ctx->ev_ready[sel], 0));
      if (b == 0) C2B_CUDA(cudaEventRecord(d0, cs));
      C2B_CUDA(cudaMemcpyAsync(ctx->h_offsets.as<uint64_t>() + c0, ctx->out_offsets[sel].p, (nc + 1) * 8,
                               cudaMemcpyDeviceToHost, cs));
      if (staged) {
        // untouched virtual memory costs nothing: reserve twice the estimate; a second reservation (after
        // draining, with the landed part carried over) only if even that was short
        for (PageBuf *pb : {&x->pg_idx, &x->pg_uv}) {
          const size_t unit = pb == &x->pg_idx ? 4 : 16;
          if (std::max<size_t>(need, 1) * unit > pb->cap) {
            drain.wait_all();
            PageBuf bigger;
            C2B_TRY(bigger.ensure_fresh(2 * std::max<size_t>(need, 1) * unit));
            if (obs_base) memcpy(bigger.p, pb->p, obs_base * unit);
            pb->release();
            *pb = bigger;
          }
        }
        if (!copy_started) {
          copy_started = true;
          t_first_copy = std::chrono::steady_clock::now();
        }
        if (O) {
          drain.submit(ctx->out_idx[sel].p, x->pg_idx.as<uint32_t>() + obs_base, O * 4, ctx->ev_ready[sel], (int)b);
          drain.submit(ctx->out_uv[sel].p, x->pg_uv.as<double>() + 2 * obs_base, O * 16, ctx->ev_ready[sel], (int)b);
        }
      } else {
        C2B_TRY(grow_pinned(ctx, ctx->h_idx, std::max<size_t>(need, 1) * 4, obs_base * 4));
        C2B_TRY(grow_pinned(ctx, ctx->h_uv, std::max<size_t>(need, 1) * 16, obs_base * 16));
        if (O) {
          C2B_CUDA(cudaMemcpyAsync(ctx->h_idx.as<uint32_t>() + obs_base, ctx->out_idx[sel].p, O * 4,
                                   cudaMemcpyDeviceToHost, cs));
          C2B_CUDA(cudaMemcpyAsync(ctx->h_uv.as<double>() + 2 * obs_base, ctx->out_uv[sel].p, O * 16,
                                   cudaMemcpyDeviceToHost, cs));
        }
        C2B_CUDA(cudaEventRecord(ctx->ev_copied[sel], cs));
      }
      obs_base += O;
      acc.n_candidates += s1.n_candidates;
      acc.pairs_evaluated += s1.pairs_evaluated;
      acc.nodes_visited += s1.nodes_visited;
      acc.tris_tested += s1.tris_tested;
      acc.ms_prep += s1.ms_prep;
      acc.ms_cull += s1.ms_cull;
      acc.ms_sort += s1.ms_sort;
      acc.ms_traverse += s1.ms_traverse;
      acc.ms_compact += s1.ms_compact;
      acc.ms_total += s1.ms_total;
    }
    C2B_CUDA(cudaEventRecord(d1, cs));
    C2B_CUDA(cudaStreamSynchronize(cs));
    C2B_CUDA(cudaStreamSynchronize(st));
    C2B_CUDA(cudaEventElapsedTime(&ms_upload, u0, u1));
    C2B_CUDA(cudaEventElapsedTime(&ms_copy, d0, d1));
    if (staged) {
      const cudaError_t e = drain.finish();
      if (e != cudaSuccess) return set_error(C2B_ERR_CUDA, "staged device-to-host copy failed: %s", cudaGetErrorString(e));
      if (copy_started)
        ms_copy = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_first_copy).count();
    }
    return C2B_OK;
  };
  rc = run();
  x->grid_locked = false;
  cudaEventDestroy(u0);
  cudaEventDestroy(u1);
  cudaEventDestroy(d0);
  cudaEventDestroy(d1);
  ctx->out_sel = 0;
  x->have_result = false;  // the resident state now holds the last batch only
  if (rc != C2B_OK) {
    cudaStreamSynchronize(cs);
    return rc;
  }
  acc.n_cameras = C;
  acc.n_obs = obs_base;
  acc.offsets = ctx->h_offsets.as<uint64_t>();
  acc.point_idx = staged ? x->pg_idx.as<uint32_t>() : ctx->h_idx.as<uint32_t>();
  acc.uv = staged ? x->pg_uv.as<double>() : ctx->h_uv.as<double>();
  acc.ms_h2d = ms_upload;
  acc.h2d_bytes = (resident_pts ? 0 : P * 24) + C * 120;
  acc.d2h_bytes = (C + 1) * 8 + obs_base * 20;
  acc.ms_d2h = ms_copy;  // first to last result copy on the copy stream; overlaps the compute of later batches
  acc.ms_total += ms_upload;
  if (C) {
    x->hist.valid = true;
    x->hist.ms_compute_per_cam = (acc.ms_prep + acc.ms_cull + acc.ms_sort + acc.ms_traverse + acc.ms_compact) / (double)C;
    x->hist.d2h_bytes_per_cam = (double)acc.d2h_bytes / (double)C;
  }
  *out = acc;
  return C2B_OK;
}

void c2b_obs_free(c2b_ctx *ctx, c2b_obs *obs) {
  (void)ctx;  // the pinned buffers are pooled in the ctx and reused by the next call
  if (!obs) return;
  obs->offsets = nullptr;
  obs->point_idx = nullptr;
  obs->uv = nullptr;
  obs->n_obs = 0;
}

int c2b_reprojection_error_resident(c2b_ctx *ctx, double norm, double *out) {
  if (!ctx || !out) return set_error(C2B_ERR_INVALID, "c2b_reprojection_error_resident: null argument");
  CtxExtra *x = extra_of(ctx);
  if (!x->have_result) return set_error(C2B_ERR_INVALID, "no resident result");
  C2B_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const uint64_t C = ctx->out_C, O = ctx->out_O, P = ctx->P;
  *out = 0.0;
  if (O == 0) return C2B_OK;
  int nb = (int)std::min<uint64_t>(blocks_for(O, ST_THREADS), ST_BLOCKS);
  C2B_TRY(ctx->misc.ensure((size_t)nb * 8 + 64));
  const double *px = ctx->pts.as<double>();
  k_reproj_partial<<<nb, ST_THREADS, 0, st>>>(ctx->cams.as<double>(), px, px + P, px + 2 * P,
                                              ctx->out_offsets[ctx->out_sel].as<uint64_t>(), C,
                                              ctx->out_idx[ctx->out_sel].as<uint32_t>(), ctx->out_uv[ctx->out_sel].as<double2>(), O,
                                              norm, ctx->misc.as<double>());
  C2B_KERNEL_CHECK();
  std::vector<double> h(nb);
  C2B_CUDA(cudaMemcpyAsync(h.data(), ctx->misc.p, (size_t)nb * 8, cudaMemcpyDeviceToHost, st));
  C2B_CUDA(cudaStreamSynchronize(st));
  double s = 0;
  for (double v : h) s += v;
  *out = std::pow(s, 1.0 / norm);
  return C2B_OK;
}

// ---- noise ----------------------------------------------------------------------------------------------
namespace {
// the arrays a noise pass works on, all on the device
struct NoiseView {
  double *cams = nullptr;     // [15 C]
  double *centers = nullptr;  // [3 C] SoA, up to date with cams
  double *pts = nullptr;      // [3 P] xyz records
  double2 *uv = nullptr;      // [O]
  uint64_t C = 0, P = 0, O = 0;
};

// device time of the last noise call: [0] upload, [1] statistics + kernels, [2] download
struct NoiseTimer {
  c2b_ctx *ctx;
  cudaEvent_t ev[4] = {};
  explicit NoiseTimer(c2b_ctx *c) : ctx(c) {
    for (auto &e : ev) cudaEventCreate(&e);
    cudaEventRecord(ev[0], ctx->stream);
  }
  void mark(int k) { cudaEventRecord(ev[k], ctx->stream); }
  void finish() {
    for (int k = 0; k < 3; ++k) {
      float t = 0;
      if (cudaEventElapsedTime(&t, ev[k], ev[k + 1]) != cudaSuccess) (void)cudaGetLastError();
      ctx->noise_ms[k] = t;
    }
  }
  ~NoiseTimer() {
    for (auto &e : ev) cudaEventDestroy(e);
  }
};

inline unsigned nz_grid(c2b_ctx *ctx, uint64_t n) {
  return (unsigned)std::min<uint64_t>(std::max<uint64_t>(blocks_for(n, NZ_THREADS), 1), (uint64_t)ctx->sm_count * 8);
}

// mean / std / nearest-origin of the chained sequence on the device.  scratch layout (doubles):
// [0..2] mean, [3..5] sumsq, [6..8] origin, then partials.
int device_stats(c2b_ctx *ctx, const NoiseView &v, double mean[3], double sd[3], bool want_origin, bool to_host = true) {
  cudaStream_t st = ctx->stream;
  const uint64_t C = v.C, P = v.P, n = C + P;
  const double num = (double)n;
  int blocks = (int)std::min<uint64_t>(std::max<uint64_t>(blocks_for(n, ST_THREADS), 1), ST_BLOCKS);
  C2B_TRY(ctx->nz_scratch.ensure((size_t)(16 + 8 * blocks) * 8));
  double *s = ctx->nz_scratch.as<double>();
  double *partial = s + 16;
  const double *cx = v.centers;
  const double *pts = v.pts;
  // the chain stays on the device: the second pass reads the mean the first one left in s[0..2]
  k_stats_partial<0><<<blocks, ST_THREADS, 0, st>>>(cx, cx + C, cx + 2 * C, C, pts, P, num, s, partial);
  C2B_KERNEL_CHECK();
  k_stats_final<<<1, 32, 0, st>>>(partial, blocks, s);
  C2B_KERNEL_CHECK();
  k_stats_partial<1><<<blocks, ST_THREADS, 0, st>>>(cx, cx + C, cx + 2 * C, C, pts, P, num, s, partial);
  C2B_KERNEL_CHECK();
  k_stats_final<<<1, 32, 0, st>>>(partial, blocks, s + 3);
  C2B_KERNEL_CHECK();
  k_bal_std<<<1, 32, 0, st>>>(s, num);
  C2B_KERNEL_CHECK();
  if (want_origin) {
    double *pd = partial;
    unsigned long long *pi = reinterpret_cast<unsigned long long *>(partial + blocks);
    k_nearest_partial<<<blocks, ST_THREADS, 0, st>>>(cx, cx + C, cx + 2 * C, C, pts, P, pd, pi);
    C2B_KERNEL_CHECK();
    k_nearest_final<<<1, 32, 0, st>>>(pd, pi, blocks, cx, cx + C, cx + 2 * C, C, pts, s + 6);
    C2B_KERNEL_CHECK();
  }
  if (to_host) {
    double h[6];
    C2B_CUDA(cudaMemcpyAsync(h, s, 48, cudaMemcpyDeviceToHost, st));
    C2B_CUDA(cudaStreamSynchronize(st));
    for (int k = 0; k < 3; ++k) {
      mean[k] = h[k];
      sd[k] = std::sqrt(h[3 + k] / num);
    }
  }
  return C2B_OK;
}

// host arrays -> the ctx's noise buffers (grow-only, kept between calls)
int noise_upload(c2b_ctx *ctx, NoiseView &v, const double *cams, uint64_t C, const double *pts, uint64_t P) {
  cudaStream_t st = ctx->stream;
  C2B_TRY(ctx->nz_cams.ensure(std::max<uint64_t>(C, 1) * 120));
  C2B_TRY(ctx->nz_centers.ensure(std::max<uint64_t>(C, 1) * 24));
  C2B_TRY(ctx->nz_pts.ensure(std::max<uint64_t>(P, 1) * 24));
  v.cams = ctx->nz_cams.as<double>();
  v.centers = ctx->nz_centers.as<double>();
  v.pts = ctx->nz_pts.as<double>();
  v.C = C;
  v.P = P;
  if (C) {
    C2B_TRY(copy_in(ctx, v.cams, cams, C * 120, cudaMemcpyHostToDevice));
    k_cam_prep<<<blocks_for(C, 128), 128, 0, st>>>(v.cams, C, v.centers, v.centers + C, v.centers + 2 * C);
    C2B_KERNEL_CHECK();
  }
  if (P) C2B_TRY(copy_in(ctx, v.pts, pts, P * 24, cudaMemcpyHostToDevice));
  return C2B_OK;
}

// the resident problem as a NoiseView (cameras + points of the last uploads, (u, v) of the last resident pass)
int resident_view(c2b_ctx *ctx, NoiseView &v, const char *who) {
  CtxExtra *x = extra_of(ctx);
  if (!x->have_points || !x->have_cameras) return set_error(C2B_ERR_INVALID, "%s: upload cameras and points first", who);
  v.cams = ctx->cams.as<double>();
  v.centers = ctx->cam_center.as<double>();
  v.pts = ctx->pts_aos.as<double>();
  v.C = ctx->C;
  v.P = ctx->P;
  v.uv = x->have_result ? ctx->out_uv[ctx->out_sel].as<double2>() : nullptr;
  v.O = x->have_result ? ctx->out_O : 0;
  return C2B_OK;
}

// after a resident noise pass: camera centres and the SoA points / bounds / grid follow the new values
int resident_refresh(c2b_ctx *ctx, bool cams_moved, bool pts_moved) {
  if (cams_moved && ctx->C) {
    double *cx = ctx->cam_center.as<double>();
    k_cam_prep<<<blocks_for(ctx->C, 128), 128, 0, ctx->stream>>>(ctx->cams.as<double>(), ctx->C, cx, cx + ctx->C, cx + 2 * ctx->C);
    C2B_KERNEL_CHECK();
  }
  if (pts_moved) {
    CtxExtra *x = extra_of(ctx);
    const bool had = x->have_result;
    C2B_TRY(c2b_points_commit(ctx, ctx->P));
    x->have_result = had;  // the CSR is still the graph of this problem
  }
  C2B_CUDA(cudaStreamSynchronize(ctx->stream));
  return C2B_OK;
}

int drift_kernels(c2b_ctx *ctx, const NoiseView &v, double strength, double angle_strength, double std_,
                  const double *dir_in, bool normalized, uint64_t seed) {
  cudaStream_t st = ctx->stream;
  double mean[3], sd[3];
  C2B_TRY(device_stats(ctx, v, mean, sd, true));
  V3 dir;
  if (normalized) {
    // add_drift_normalized, src/noise.rs:53-55
    V3 s{sd[0], sd[1], sd[2]};
    double m = std::sqrt((s.x * s.x + s.y * s.y) + s.z * s.z);
    double inv = 1.0 / m;
    dir = V3{s.x * inv, s.y * inv, s.z * inv};
    strength = strength * m;
  } else {
    dir = V3{dir_in[0], dir_in[1], dir_in[2]};
  }
  const double *origin = ctx->nz_scratch.as<double>() + 6;
  if (v.C) {
    k_drift_cams<<<blocks_for(v.C, NZ_THREADS), NZ_THREADS, 0, st>>>(v.cams, v.C, origin, dir, strength, angle_strength, std_, seed);
    C2B_KERNEL_CHECK();
  }
  if (v.P) {
    k_drift_pts<<<nz_grid(ctx, v.P), NZ_THREADS, 0, st>>>(v.pts, v.P, origin, dir, strength, std_, seed);
    C2B_KERNEL_CHECK();
  }
  return C2B_OK;
}

int drift_impl(c2b_ctx *ctx, double *cams, uint64_t C, double *pts, uint64_t P, double strength,
               double angle_strength, double std_, const double *dir_in, bool normalized,
               uint64_t seed) {
  if (!ctx || (C && !cams) || (P && !pts)) return set_error(C2B_ERR_INVALID, "add_drift: null argument");
  if (C + P == 0) return set_error(C2B_ERR_EMPTY, "add_drift: problem has no cameras and no points");
  C2B_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  NoiseView v;
  NoiseTimer tm(ctx);
  C2B_TRY(noise_upload(ctx, v, cams, C, pts, P));
  tm.mark(1);
  C2B_TRY(drift_kernels(ctx, v, strength, angle_strength, std_, dir_in, normalized, seed));
  tm.mark(2);
  if (C) C2B_CUDA(cudaMemcpyAsync(cams, v.cams, C * 120, cudaMemcpyDeviceToHost, st));
  if (P) C2B_CUDA(cudaMemcpyAsync(pts, v.pts, P * 24, cudaMemcpyDeviceToHost, st));
  tm.mark(3);
  C2B_CUDA(cudaStreamSynchronize(st));
  tm.finish();
  return C2B_OK;
}

// cameras and points of add_noise (the observations are separate: the host entry streams them in chunks)
int noise_kernels(c2b_ctx *ctx, const NoiseView &v, double translation_std, double rotation_std, double point_std,
                  uint64_t seed) {
  cudaStream_t st = ctx->stream;
  // |std()| scales the camera translations (src/noise.rs:127,141); it stays on the device (scratch[9])
  C2B_TRY(device_stats(ctx, v, nullptr, nullptr, false, false));
  if (v.C) {
    k_noise_cams<<<blocks_for(v.C, NZ_THREADS), NZ_THREADS, 0, st>>>(v.cams, v.C, ctx->nz_scratch.as<double>() + 9,
                                                                    translation_std, rotation_std, seed);
    C2B_KERNEL_CHECK();
  }
  // a zero standard deviation adds unit_random() * 0 (src/noise.rs:149,159-168): the array is unchanged, so its
  // kernel — and in the host entry its round trip over PCIe — is skipped
  if (v.P && point_std != 0.0) {
    k_noise_pts<<<nz_grid(ctx, v.P), NZ_THREADS, 0, st>>>(v.pts, v.P, point_std, seed);
    C2B_KERNEL_CHECK();
  }
  return C2B_OK;
}

int sin_kernels(c2b_ctx *ctx, const NoiseView &v, const double dir[3], const double noise_dir[3], double strength,
                double frequency) {
  cudaStream_t st = ctx->stream;
  const uint64_t C = v.C, P = v.P, n = C + P;
  int blocks = (int)std::min<uint64_t>(std::max<uint64_t>(blocks_for(n, ST_THREADS), 1), ST_BLOCKS);
  C2B_TRY(ctx->nz_scratch.ensure((size_t)(8 + 6 * blocks) * 8));
  double *s = ctx->nz_scratch.as<double>();
  const double *cx = v.centers;
  k_extent_partial<<<blocks, ST_THREADS, 0, st>>>(cx, cx + C, cx + 2 * C, C, v.pts, P, s + 8);
  C2B_KERNEL_CHECK();
  k_extent_final<<<1, 32, 0, st>>>(s + 8, blocks, s);
  C2B_KERNEL_CHECK();
  double ext[6];
  C2B_CUDA(cudaMemcpyAsync(ext, s, 48, cudaMemcpyDeviceToHost, st));
  C2B_CUDA(cudaStreamSynchronize(st));
  V3 dim{ext[3] - ext[0], ext[4] - ext[1], ext[5] - ext[2]};
  // "Add epsilon to nonexistent dimensions" (src/noise.rs:395-396)
  if (dim.x == 0.0) dim.x = 1e-8;
  if (dim.y == 0.0) dim.y = 1e-8;
  if (dim.z == 0.0) dim.z = 1e-8;
  const double nm = std::sqrt((noise_dir[0] * noise_dir[0] + noise_dir[1] * noise_dir[1]) + noise_dir[2] * noise_dir[2]);
  const double inv = 1.0 / nm;
  const V3 nd{noise_dir[0] * inv, noise_dir[1] * inv, noise_dir[2] * inv}, d{dir[0], dir[1], dir[2]};
  if (C) {
    k_sin_cams<<<blocks_for(C, 128), 128, 0, st>>>(v.cams, C, dim, d, nd, strength, frequency);
    C2B_KERNEL_CHECK();
  }
  if (P) {
    k_sin_pts<<<blocks_for(P, 256), 256, 0, st>>>(v.pts, P, dim, d, nd, strength, frequency);
    C2B_KERNEL_CHECK();
  }
  return C2B_OK;
}
}  // namespace

int c2b_add_drift(c2b_ctx *ctx, double *cams, uint64_t C, double *pts, uint64_t P, double strength,
                  double angle_strength, double std_, const double dir[3], uint64_t seed) {
  if (!dir) return set_error(C2B_ERR_INVALID, "c2b_add_drift: dir is null");
  return drift_impl(ctx, cams, C, pts, P, strength, angle_strength, std_, dir, false, seed);
}

int c2b_add_drift_normalized(c2b_ctx *ctx, double *cams, uint64_t C, double *pts, uint64_t P,
                             double strength, double angle_strength, double std_, uint64_t seed) {
  return drift_impl(ctx, cams, C, pts, P, strength, angle_strength, std_, nullptr, true, seed);
}

int c2b_add_noise(c2b_ctx *ctx, double *cams, uint64_t C, double *pts, uint64_t P, double *uv,
                  uint64_t O, double translation_std, double rotation_std, double point_std,
                  double observations_std, uint64_t seed) {
  if (!ctx || (C && !cams) || (P && !pts) || (O && !uv))
    return set_error(C2B_ERR_INVALID, "c2b_add_noise: null argument");
  if (C + P == 0) return set_error(C2B_ERR_EMPTY, "add_noise: problem has no cameras and no points");
  C2B_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream, cs = ctx->copy_stream;
  NoiseView v;
  NoiseTimer tm(ctx);
  if (observations_std == 0.0) O = 0;
  // The observations (16 B each way, by far the largest array) stream through the GPU in chunks on the copy
  // stream while cameras and points are handled on the main one: PCIe is full duplex, so chunk k's download
  // overlaps chunk k+1's upload and the call costs about one direction of the transfer, not two.
  cudaEvent_t obs_done = nullptr;
  if (O) {
    const uint64_t chunk = std::max<uint64_t>(1u << 20, (O + 15) / 16);
    C2B_TRY(ctx->nz_uv.ensure(std::min(O, 2 * chunk) * 16));
    double2 *buf = ctx->nz_uv.as<double2>();
    cudaEvent_t slot_free[2] = {nullptr, nullptr}, up[2] = {nullptr, nullptr};
    for (int k = 0; k < 2; ++k) {
      C2B_CUDA(cudaEventCreateWithFlags(&slot_free[k], cudaEventDisableTiming));
      C2B_CUDA(cudaEventCreateWithFlags(&up[k], cudaEventDisableTiming));
    }
    C2B_CUDA(cudaEventCreateWithFlags(&obs_done, cudaEventDisableTiming));
    // uploads + kernels on the aux stream, downloads on the copy stream; two slots
    cudaStream_t ax = ctx->aux_stream;
    uint64_t k = 0;
    for (uint64_t first = 0; first < O; first += chunk, ++k) {
      const uint64_t n = std::min(chunk, O - first);
      const int slot = (int)(k & 1);
      double2 *d = buf + (uint64_t)slot * chunk;
      if (k >= 2) C2B_CUDA(cudaStreamWaitEvent(ax, slot_free[slot], 0));
      C2B_CUDA(cudaMemcpyAsync(d, uv + 2 * first, n * 16, cudaMemcpyHostToDevice, ax));
      k_noise_obs<<<nz_grid(ctx, n), NZ_THREADS, 0, ax>>>(d, n, first, observations_std, seed);
      C2B_KERNEL_CHECK();
      C2B_CUDA(cudaEventRecord(up[slot], ax));
      C2B_CUDA(cudaStreamWaitEvent(cs, up[slot], 0));
      C2B_CUDA(cudaMemcpyAsync(uv + 2 * first, d, n * 16, cudaMemcpyDeviceToHost, cs));
      C2B_CUDA(cudaEventRecord(slot_free[slot], cs));
    }
    C2B_CUDA(cudaEventRecord(obs_done, cs));
    for (int q = 0; q < 2; ++q) {
      cudaEventDestroy(slot_free[q]);
      cudaEventDestroy(up[q]);
    }
  }
  C2B_TRY(noise_upload(ctx, v, cams, C, pts, P));
  tm.mark(1);
  C2B_TRY(noise_kernels(ctx, v, translation_std, rotation_std, point_std, seed));
  tm.mark(2);
  if (C) C2B_CUDA(cudaMemcpyAsync(cams, v.cams, C * 120, cudaMemcpyDeviceToHost, st));
  if (P && point_std != 0.0) C2B_CUDA(cudaMemcpyAsync(pts, v.pts, P * 24, cudaMemcpyDeviceToHost, st));
  if (obs_done) {
    C2B_CUDA(cudaStreamWaitEvent(st, obs_done, 0));
    cudaEventDestroy(obs_done);
  }
  tm.mark(3);
  C2B_CUDA(cudaStreamSynchronize(st));
  tm.finish();
  return C2B_OK;
}

int c2b_add_sin_noise(c2b_ctx *ctx, double *cams, uint64_t C, double *pts, uint64_t P, const double dir[3],
                      const double noise_dir[3], double strength, double frequency) {
  if (!ctx || (C && !cams) || (P && !pts) || !dir || !noise_dir)
    return set_error(C2B_ERR_INVALID, "c2b_add_sin_noise: null argument");
  if (C + P == 0) return set_error(C2B_ERR_EMPTY, "add_sin_noise: problem has no cameras and no points");
  C2B_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  NoiseView v;
  NoiseTimer tm(ctx);
  C2B_TRY(noise_upload(ctx, v, cams, C, pts, P));
  tm.mark(1);
  C2B_TRY(sin_kernels(ctx, v, dir, noise_dir, strength, frequency));
  tm.mark(2);
  if (C) C2B_CUDA(cudaMemcpyAsync(cams, v.cams, C * 120, cudaMemcpyDeviceToHost, st));
  if (P) C2B_CUDA(cudaMemcpyAsync(pts, v.pts, P * 24, cudaMemcpyDeviceToHost, st));
  tm.mark(3);
  C2B_CUDA(cudaStreamSynchronize(st));
  tm.finish();
  return C2B_OK;
}

// ---- the same passes on the RESIDENT problem: no PCIe traffic at all -------------------------------------
int c2b_add_drift_resident(c2b_ctx *ctx, double strength, double angle_strength, double std_, const double *dir,
                           uint64_t seed) {
  if (!ctx) return set_error(C2B_ERR_INVALID, "c2b_add_drift_resident: null ctx");
  C2B_CUDA(cudaSetDevice(ctx->device));
  NoiseView v;
  C2B_TRY(resident_view(ctx, v, "c2b_add_drift_resident"));
  if (v.C + v.P == 0) return set_error(C2B_ERR_EMPTY, "add_drift: problem has no cameras and no points");
  NoiseTimer tm(ctx);
  tm.mark(1);
  C2B_TRY(drift_kernels(ctx, v, strength, angle_strength, std_, dir, dir == nullptr, seed));
  tm.mark(2);
  tm.mark(3);
  C2B_TRY(resident_refresh(ctx, true, true));
  tm.finish();
  return C2B_OK;
}

int c2b_add_noise_resident(c2b_ctx *ctx, double translation_std, double rotation_std, double point_std,
                           double observations_std, uint64_t seed) {
  if (!ctx) return set_error(C2B_ERR_INVALID, "c2b_add_noise_resident: null ctx");
  C2B_CUDA(cudaSetDevice(ctx->device));
  NoiseView v;
  C2B_TRY(resident_view(ctx, v, "c2b_add_noise_resident"));
  if (v.C + v.P == 0) return set_error(C2B_ERR_EMPTY, "add_noise: problem has no cameras and no points");
  NoiseTimer tm(ctx);
  tm.mark(1);
  C2B_TRY(noise_kernels(ctx, v, translation_std, rotation_std, point_std, seed));
  if (v.O && observations_std != 0.0) {
    k_noise_obs<<<nz_grid(ctx, v.O), NZ_THREADS, 0, ctx->stream>>>(v.uv, v.O, 0, observations_std, seed);
    C2B_KERNEL_CHECK();
  }
  tm.mark(2);
  tm.mark(3);
  C2B_TRY(resident_refresh(ctx, true, point_std != 0.0));
  tm.finish();
  return C2B_OK;
}

int c2b_add_sin_noise_resident(c2b_ctx *ctx, const double dir[3], const double noise_dir[3], double strength,
                               double frequency) {
  if (!ctx || !dir || !noise_dir) return set_error(C2B_ERR_INVALID, "c2b_add_sin_noise_resident: null argument");
  C2B_CUDA(cudaSetDevice(ctx->device));
  NoiseView v;
  C2B_TRY(resident_view(ctx, v, "c2b_add_sin_noise_resident"));
  if (v.C + v.P == 0) return set_error(C2B_ERR_EMPTY, "add_sin_noise: problem has no cameras and no points");
  NoiseTimer tm(ctx);
  tm.mark(1);
  C2B_TRY(sin_kernels(ctx, v, dir, noise_dir, strength, frequency));
  tm.mark(2);
  tm.mark(3);
  C2B_TRY(resident_refresh(ctx, true, true));
  tm.finish();
  return C2B_OK;
}

int c2b_download_problem(c2b_ctx *ctx, double *cams_out, double *pts_out) {
  if (!ctx) return set_error(C2B_ERR_INVALID, "c2b_download_problem: null ctx");
  NoiseView v;
  C2B_TRY(resident_view(ctx, v, "c2b_download_problem"));
  if ((v.C && !cams_out) || (v.P && !pts_out)) return set_error(C2B_ERR_INVALID, "c2b_download_problem: null output array");
  C2B_CUDA(cudaSetDevice(ctx->device));
  if (v.C) C2B_CUDA(cudaMemcpyAsync(cams_out, v.cams, v.C * 120, cudaMemcpyDeviceToHost, ctx->stream));
  if (v.P) C2B_CUDA(cudaMemcpyAsync(pts_out, v.pts, v.P * 24, cudaMemcpyDeviceToHost, ctx->stream));
  C2B_CUDA(cudaStreamSynchronize(ctx->stream));
  return C2B_OK;
}

int c2b_noise_timing(c2b_ctx *ctx, float ms[3]) {
  if (!ctx || !ms) return set_error(C2B_ERR_INVALID, "c2b_noise_timing: null argument");
  for (int k = 0; k < 3; ++k) ms[k] = ctx->noise_ms[k];
  return C2B_OK;
}

// ---- input generation: world points (src/generate.rs:356-420) -----------------------------------------
int c2b_generate_world_points_uniform(c2b_ctx *ctx, const float *xyz, uint64_t nv, const uint32_t *tri,
                                      uint64_t nt, const double *cams, uint64_t C, uint64_t num_points,
                                      double max_dist, uint64_t seed, double *pts_out, uint64_t *n_out) {
  if (!ctx || !n_out || (nv && !xyz) || (nt && !tri) || (C && !cams) || (num_points && !pts_out))
    return set_error(C2B_ERR_INVALID, "c2b_generate_world_points_uniform: null argument");
  *n_out = 0;
  if (C == 0)
    return set_error(C2B_ERR_INVALID, "Cannot generate world points with 0 cameras. Try increasing the number of "
                                      "cameras generated (via --cameras).");
  if (num_points == 0) return C2B_OK;
  if (num_points >= 0x7fffffffull) return set_error(C2B_ERR_INVALID, "too many points requested");
  C2B_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  // areas and their running sum (sequential f64, the WeightedIndex table); vertices as float4 triples
  std::vector<double> cdf(nt);
  std::vector<float4> tv(3 * nt);
  double run = 0.0;
  for (uint64_t i = 0; i < nt; ++i) {
    V3 v[3];
    for (int j = 0; j < 3; ++j) {
      const uint64_t k = tri[3 * i + j];
      if (k >= nv) return set_error(C2B_ERR_INVALID, "triangle %llu references vertex >= %llu", (unsigned long long)i,
                                    (unsigned long long)nv);
      tv[3 * i + j] = make_float4(xyz[3 * k], xyz[3 * k + 1], xyz[3 * k + 2], 0.0f);
      v[j] = V3{(double)xyz[3 * k], (double)xyz[3 * k + 1], (double)xyz[3 * k + 2]};
    }
    const V3 e1{v[1].x - v[0].x, v[1].y - v[0].y, v[1].z - v[0].z}, e2{v[2].x - v[0].x, v[2].y - v[0].y, v[2].z - v[0].z};
    run += mag(cross(e1, e2)) / 2.0;
    cdf[i] = run;
  }
  if (nt == 0 || !(run > 0.0) || !std::isfinite(run))
    return set_error(C2B_ERR_INVALID, "mesh has no triangle of positive area to sample points on");
  // camera centres -> uniform grid with cells of side >= max_dist
  std::vector<double> cen(3 * C);
  double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (uint64_t i = 0; i < C; ++i) {
    const V3 c = camera_center(cams + 15 * i);
    const double a[3] = {c.x, c.y, c.z};
    for (int k = 0; k < 3; ++k) {
      cen[3 * i + k] = a[k];
      if (a[k] < lo[k]) lo[k] = a[k];
      if (a[k] > hi[k]) hi[k] = a[k];
    }
  }
  CamGrid g;
  double h = (max_dist > 0.0 && std::isfinite(max_dist)) ? max_dist : 1.0;
  for (int k = 0; k < 3; ++k) {
    if (!std::isfinite(lo[k]) || !std::isfinite(hi[k])) lo[k] = hi[k] = 0.0;  // NaN / inf centres: one cell
    h = std::max(h, (hi[k] - lo[k]) / 128.0);
  }
  if (!std::isfinite(max_dist) && max_dist > 0.0) h = INFINITY;  // everything is near: a single cell
  uint64_t cells = 1;
  for (int k = 0; k < 3; ++k) {
    g.lo[k] = lo[k];
    g.n[k] = std::isfinite(h) ? (int)std::floor((hi[k] - lo[k]) / h) + 1 : 1;
    cells *= (uint64_t)g.n[k];
  }
  g.inv_h = std::isfinite(h) ? 1.0 / h : 0.0;
  std::vector<uint32_t> cell_start(cells + 1, 0), cell_of(C);
  for (uint64_t i = 0; i < C; ++i) {
    const int cx = cam_grid_coord(g, 0, cen[3 * i]), cy = cam_grid_coord(g, 1, cen[3 * i + 1]),
              cz = cam_grid_coord(g, 2, cen[3 * i + 2]);
    cell_of[i] = ((uint32_t)cz * g.n[1] + cy) * g.n[0] + cx;
    cell_start[cell_of[i] + 1]++;
  }
  for (uint64_t c = 0; c < cells; ++c) cell_start[c + 1] += cell_start[c];
  std::vector<double> cen_sorted(3 * C);
  {
    std::vector<uint32_t> cur(cell_start.begin(), cell_start.end() - 1);
    for (uint64_t i = 0; i < C; ++i) {
      const uint32_t p = cur[cell_of[i]]++;
      for (int k = 0; k < 3; ++k) cen_sorted[3 * p + k] = cen[3 * i + k];
    }
  }
  DevBuf d_tv, d_cdf, d_cs, d_cen, d_cand, d_keep, d_rank, d_out, d_small;
  auto cleanup = [&]() {
    for (DevBuf *b : {&d_tv, &d_cdf, &d_cs, &d_cen, &d_cand, &d_keep, &d_rank, &d_out, &d_small}) b->release();
  };
  int rc = [&]() -> int {
    C2B_TRY(d_tv.ensure(tv.size() * 16));
    C2B_TRY(d_cdf.ensure(nt * 8));
    C2B_TRY(d_cs.ensure((cells + 1) * 4));
    C2B_TRY(d_cen.ensure(C * 24));
    C2B_TRY(d_out.ensure(num_points * 24));
    C2B_TRY(d_small.ensure(64));
    C2B_CUDA(cudaMemcpyAsync(d_tv.p, tv.data(), tv.size() * 16, cudaMemcpyHostToDevice, st));
    C2B_CUDA(cudaMemcpyAsync(d_cdf.p, cdf.data(), nt * 8, cudaMemcpyHostToDevice, st));
    C2B_CUDA(cudaMemcpyAsync(d_cs.p, cell_start.data(), (cells + 1) * 4, cudaMemcpyHostToDevice, st));
    C2B_CUDA(cudaMemcpyAsync(d_cen.p, cen_sorted.data(), C * 24, cudaMemcpyHostToDevice, st));
    SampleArgs a;
    a.tri_v = d_tv.as<float4>();
    a.cdf = d_cdf.as<double>();
    a.nt = nt;
    a.total = run;
    a.g = g;
    a.cell_start = d_cs.as<uint32_t>();
    a.cen = d_cen.as<double>();
    a.max_d2 = max_dist * max_dist;
    a.seed = seed;
    const uint64_t threshold = 10 * num_points;
    uint64_t got = 0, consumed = 0;
    while (got < num_points) {
      const uint64_t want = num_points - got;
      const uint64_t n = std::min<uint64_t>(want + want / 4 + 1024, 1ull << 26);
      C2B_TRY(d_cand.ensure(n * 24));
      C2B_TRY(d_keep.ensure((n + 1) * 4));
      C2B_TRY(d_rank.ensure((n + 1) * 4));
      a.first = consumed;
      a.n = n;
      a.cand = d_cand.as<double>();
      a.keep = d_keep.as<uint32_t>();
      C2B_CUDA(cudaMemsetAsync(d_small.p, 0, 64, st));
      k_sample_points<<<blocks_for(n, 256), 256, 0, st>>>(a);
      C2B_KERNEL_CHECK();
      uint32_t *d_total = d_small.as<uint32_t>() + 4;
      C2B_TRY(exclusive_scan_u32(st, a.keep, d_rank.as<uint32_t>(), n, d_total, ctx->scan_tmp));
      k_sample_compact<<<blocks_for(n, 256), 256, 0, st>>>(a.cand, a.keep, d_rank.as<uint32_t>(), n, got, num_points,
                                                           d_out.as<double>(), d_small.as<unsigned long long>());
      C2B_KERNEL_CHECK();
      unsigned long long h[4];
      C2B_CUDA(cudaMemcpyAsync(h, d_small.p, 32, cudaMemcpyDeviceToHost, st));
      C2B_CUDA(cudaStreamSynchronize(st));
      uint32_t accepted;
      memcpy(&accepted, reinterpret_cast<const char *>(h) + 16, 4);
      if (got + accepted >= num_points) {
        // h[0] = round-local index of the candidate that completed the request
        const uint64_t used = consumed + h[0] + 1;
        if (used - num_points >= threshold)
          return set_error(C2B_ERR_INVALID, "Failed to generate enough points. %llu successes, %llu failures, %llu "
                                            "requested points.", (unsigned long long)num_points,
                           (unsigned long long)(used - num_points), (unsigned long long)num_points);
        got = num_points;
        break;
      }
      got += accepted;
      consumed += n;
      if (consumed - got >= threshold)
        return set_error(C2B_ERR_INVALID, "Failed to generate enough points. %llu successes, %llu failures, %llu "
                                          "requested points.", (unsigned long long)got,
                         (unsigned long long)(consumed - got), (unsigned long long)num_points);
    }
    C2B_CUDA(cudaMemcpyAsync(pts_out, d_out.p, num_points * 24, cudaMemcpyDeviceToHost, st));
    C2B_CUDA(cudaStreamSynchronize(st));
    return C2B_OK;
  }();
  cleanup();
  if (rc == C2B_OK) *n_out = num_points;
  return rc;
}

int c2b_probe_fp64(c2b_ctx *ctx, double *dfma_per_s) {
  if (!ctx || !dfma_per_s) return set_error(C2B_ERR_INVALID, "c2b_probe_fp64: null argument");
  C2B_CUDA(cudaSetDevice(ctx->device));
  C2B_TRY(ctx->misc.ensure(8));
  cudaStream_t st = ctx->stream;
  const int blocks = ctx->sm_count * 8, threads = 256, iters = 4096;
  cudaEvent_t e0, e1;
  C2B_CUDA(cudaEventCreate(&e0));
  C2B_CUDA(cudaEventCreate(&e1));
  float best = 0.0f;
  for (int rep = 0; rep < 4; ++rep) {  // first repetition warms up
    C2B_CUDA(cudaEventRecord(e0, st));
    k_probe_dfma<<<blocks, threads, 0, st>>>(ctx->misc.as<double>(), iters, 1.0000001);
    C2B_KERNEL_CHECK();
    C2B_CUDA(cudaEventRecord(e1, st));
    C2B_CUDA(cudaEventSynchronize(e1));
    float ms = 0.0f;
    C2B_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (rep && (best == 0.0f || ms < best)) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *dfma_per_s = (double)blocks * threads * iters * PROBE_CHAINS / (best * 1e-3);
  return C2B_OK;
}

int c2b_mean_std(c2b_ctx *ctx, const double *cams, uint64_t C, const double *pts, uint64_t P,
                 double mean[3], double sd[3]) {
  if (!ctx || (C && !cams) || (P && !pts) || !mean || !sd)
    return set_error(C2B_ERR_INVALID, "c2b_mean_std: null argument");
  if (C + P == 0) return set_error(C2B_ERR_EMPTY, "mean/std of an empty problem");
  C2B_CUDA(cudaSetDevice(ctx->device));
  NoiseView v;
  C2B_TRY(noise_upload(ctx, v, cams, C, pts, P));
  return device_stats(ctx, v, mean, sd, false);
}

}  // extern "C"
