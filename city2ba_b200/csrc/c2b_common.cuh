// c2b_common.cuh — context, error plumbing, grow-only device buffers.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#if defined(__linux__)
#include <sys/syscall.h>
#include <unistd.h>
#endif

#include "../../include/city2ba_cuda.h"

namespace c2b {

// thread-local last error (c2b_last_error)
inline std::string &last_error_ref() {
  static thread_local std::string s;
  return s;
}
inline int set_error(int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  last_error_ref() = buf;
  return code;
}

#define C2B_CUDA(call)                                                                        \
  do {                                                                                        \
    cudaError_t _e = (call);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      return c2b::set_error(_e == cudaErrorMemoryAllocation ? C2B_ERR_OOM : C2B_ERR_CUDA,     \
                            "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
    }                                                                                         \
  } while (0)

#define C2B_TRY(call)        \
  do {                       \
    int _s = (call);         \
    if (_s != C2B_OK) return _s; \
  } while (0)

// every kernel launch in the library is followed by this macro: it also counts launches
// (c2b_kernel_launches, used by bench.py's gpu_launches)
inline std::atomic<uint64_t> &launch_counter() {
  static std::atomic<uint64_t> n{0};
  return n;
}
#define C2B_KERNEL_CHECK()           \
  do {                               \
    ++c2b::launch_counter();         \
    C2B_CUDA(cudaGetLastError());    \
  } while (0)

// grow-only device buffer
struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return C2B_OK;
    if (p) {
      cudaFree(p);
      p = nullptr;
      cap = 0;
    }
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      (void)cudaGetLastError();
      e = cudaMalloc(&p, bytes);
      want = bytes;
    }
    if (e != cudaSuccess) {
      (void)cudaGetLastError();
      p = nullptr;
      return set_error(C2B_ERR_OOM, "cudaMalloc(%zu bytes) failed: %s", bytes,
                       cudaGetErrorString(e));
    }
    cap = want;
    return C2B_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <class T>
  T *as() const {
    return reinterpret_cast<T *>(p);
  }
};

// grow-only pinned host buffer
// Pinned host buffers are allocated on the NUMA node the GPU hangs off (c2b_init reads it from sysfs):
// with one process per GPU on a two-socket host, result slabs that land on the other socket cross the
// inter-socket link on their way out of every GPU at once.  Done by setting the calling thread's memory
// policy to MPOL_PREFERRED(node) around the allocation only; C2B_NUMA_LOCAL=0 switches it off.
struct NumaPreferred {
  bool active = false;
  explicit NumaPreferred(int node) {
#if defined(__linux__)
    if (node < 0 || node >= 1024) return;
    unsigned long mask[16] = {0};
    mask[node / (8 * sizeof(unsigned long))] |= 1ul << (node % (8 * sizeof(unsigned long)));
    active = syscall(SYS_set_mempolicy, 1 /* MPOL_PREFERRED */, mask, 1025ul) == 0;
#else
    (void)node;
#endif
  }
  ~NumaPreferred() {
#if defined(__linux__)
    if (active) syscall(SYS_set_mempolicy, 0 /* MPOL_DEFAULT */, nullptr, 0ul);
#endif
  }
};

struct PinBuf {
  void *p = nullptr;
  size_t cap = 0;
  int node = -1;  // NUMA node of the GPU these buffers feed (-1: unknown / off)
  bool portable = false;  // every device of the process may DMA into it (the one host CSR of c2b_visibility_graph_multi)
  int ensure(size_t bytes) {
    if (bytes <= cap) return C2B_OK;
    if (p) {
      cudaFreeHost(p);
      p = nullptr;
      cap = 0;
    }
    // pinning costs ~0.5 s per GB on this pool's hosts (profiles/r02b_alloc_probe.txt): large buffers get
    // little slack, small ones a quarter
    size_t want = bytes + (bytes > (64u << 20) ? bytes / 32 : bytes / 4) + 256;
    NumaPreferred local(node);
    cudaError_t e = cudaHostAlloc(&p, want, portable ? cudaHostAllocPortable : cudaHostAllocDefault);
    if (e != cudaSuccess) {
      (void)cudaGetLastError();
      p = nullptr;
      return set_error(C2B_ERR_OOM, "cudaHostAlloc(%zu bytes) failed: %s", want,
                       cudaGetErrorString(e));
    }
    cap = want;
    return C2B_OK;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
  }
  template <class T>
  T *as() const {
    return reinterpret_cast<T *>(p);
  }
};

enum StageEvent {
  EV_START = 0,
  EV_H2D,
  EV_PREP,
  EV_CULL,
  EV_SORT,
  EV_TRAVERSE,
  EV_COMPACT,
  EV_D2H,
  EV_COUNT
};

}  // namespace c2b

struct c2b_scene {
  int device = 0;  // the scene may outlive the ctx that built it; keep the ordinal, not the ctx
  uint64_t n_tris = 0;   // after dropping degenerate index triples
  uint64_t n_nodes = 0;  // 2*n_tris - 1 (0 when empty)
  float lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
  c2b::DevBuf nodes;  // float4[2*n_nodes]: {lo.xyz, escape}, {hi.xyz, leaf tri slot or -1}
  c2b::DevBuf tris;   // float4[3*n_tris]: v0, v1, v2 in leaf (Morton) order
};

struct c2b_ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[c2b::EV_COUNT] = {};

  // resident problem
  uint64_t C = 0, P = 0;
  c2b::DevBuf cams_all;      // host-buffer call: all of the call's cameras (its batches are slices of it)
  c2b::DevBuf cams;          // double[15*C] as given
  c2b::DevBuf cam_center;    // double[3*C]  SoA: x[C], y[C], z[C]
  c2b::DevBuf pts;           // double[3*P]  SoA: x[P], y[P], z[P] (streamed by the cull kernels)
  c2b::DevBuf pts_aos;       // double[3*P]  as given (24 B records for the per-candidate gathers)
  c2b::DevBuf stage;         // staging for H2D of AoS inputs
  c2b::PinBuf pin_in;        // pinned staging for pageable callers

  // grid over points
  c2b::DevBuf cell_of_pt, cell_start, cell_cursor, grid_x, grid_y, grid_z, grid_idx;

  // candidate pool + sort
  c2b::DevBuf pool_key, pool_uv, cam_count, counters;
  c2b::DevBuf sort_keys[2], sort_vals[2], sort_hist;
  c2b::DevBuf scan_tmp[3];
  uint64_t pool_capacity = 0;

  // traversal + compaction
  c2b::DevBuf vis_words, word_prefix;
  // device CSR, double-buffered so that the D2H copy of one camera batch overlaps the next batch
  c2b::DevBuf out_offsets[2], out_idx[2], out_uv[2];
  int out_sel = 0;
  uint64_t out_C = 0, out_O = 0;
  cudaStream_t copy_stream = nullptr;
  cudaStream_t aux_stream = nullptr;  // leaf lists beside the plan (visibility_grid_fused)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  cudaEvent_t ev_ready[2] = {}, ev_copied[2] = {};
  c2b::DevBuf misc;  // small scratch (reductions)
  c2b::DevBuf tri_list, tri_count;  // per-camera leaf lists for list-driven traversal
  // fused grid schedule: per-camera plan (row points -> scratch slice), visible counts, CSR offsets
  c2b::DevBuf ev_off, vis_count, seg_off, scratch_idx, plan_rows, plan_row_count;
  c2b::DevBuf epi_status, epi_prefix;  // in-kernel epilogue: per-camera count flags and CSR offsets
  int numa_node = -1;  // of the device, from sysfs (-1: unknown)

  // noise passes on host arrays: grow-only device copies kept between calls
  c2b::DevBuf nz_cams, nz_centers, nz_pts, nz_uv, nz_scratch;
  float noise_ms[3] = {0, 0, 0};  // last noise call: upload, statistics + kernels, download

  // host results
  c2b::PinBuf h_offsets, h_idx, h_uv, h_small;
};
