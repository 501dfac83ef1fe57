// c2b_compact.cuh — kernel (3): stream compaction of the surviving candidates into the CSR
// observation lists (src/generate.rs:473-478), plus the analytic 2-D wall occlusion of
// `city2ba synthetic` (src/synthetic.rs:52-124) which produces the same visibility words the
// ray traversal does.
//
// The traversal kernel leaves one 32-bit ballot word per warp (bit = candidate visible).  Here:
// popcount per word -> device-wide exclusive scan (warp shuffles + block scan, c2b_sort.cuh) ->
// every visible candidate writes (point index, u, v) at word_prefix + popc(lower bits), which
// keeps the sorted (camera-major, ascending point) order; CSR offsets are read off the same
// prefix at each camera's first candidate.
#pragma once
#include "c2b_common.cuh"
#include "c2b_math.cuh"

namespace c2b {

__global__ void k_word_popc(const uint32_t *__restrict__ words, uint64_t n_words,
                            uint32_t *__restrict__ counts) {
  uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w < n_words) counts[w] = __popc(words[w]);
}

__global__ void k_compact_write(const uint32_t *__restrict__ words,
                                const uint32_t *__restrict__ word_prefix,
                                const uint64_t *__restrict__ keys, const uint32_t *__restrict__ vals,
                                const double2 *__restrict__ pool_uv, uint64_t n_cand, int pbits,
                                uint32_t *__restrict__ out_idx, double2 *__restrict__ out_uv) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_cand) return;
  uint32_t word = words[i >> 5];
  uint32_t bit = (uint32_t)(i & 31);
  if (!((word >> bit) & 1u)) return;
  uint64_t pos = (uint64_t)word_prefix[i >> 5] + __popc(word & ((1u << bit) - 1u));
  out_idx[pos] = (uint32_t)(keys[i] & ((1ull << pbits) - 1ull));
  out_uv[pos] = pool_uv[vals[i]];
}

// offsets[c] = number of visible candidates before camera c's first candidate
__global__ void k_csr_offsets(const uint32_t *__restrict__ cand_start /*C+1*/, uint64_t C,
                              const uint32_t *__restrict__ words,
                              const uint32_t *__restrict__ word_prefix, uint64_t n_cand,
                              const uint32_t *__restrict__ total, uint64_t *__restrict__ offsets) {
  uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c > C) return;
  uint64_t s = cand_start[c];
  if (s >= n_cand) {
    offsets[c] = *total;
    return;
  }
  uint32_t word = words[s >> 5];
  uint32_t bit = (uint32_t)(s & 31);
  offsets[c] = (uint64_t)word_prefix[s >> 5] + __popc(word & ((1u << bit) - 1u));
}

// all candidates visible (C2B_OCC_NONE, and scenes without triangles); padding slots stay clear
__global__ void k_words_all_visible(const uint64_t *__restrict__ keys, uint32_t *__restrict__ words,
                                    uint64_t n_cand) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if ((i & ~31ull) >= n_cand) return;
  const bool vis = i < n_cand && keys[i] != ~0ull;
  const unsigned m = __ballot_sync(0xffffffffu, vis);
  if ((threadIdx.x & 31) == 0) words[i >> 5] = m;
}

// ---- FP64 issue-rate probe (c2b_probe_fp64): PROBE_CHAINS independent DFMA chains per thread ---------
constexpr int PROBE_CHAINS = 8;
__global__ void __launch_bounds__(256) k_probe_dfma(double *__restrict__ sink, int iters, double m) {
  double a[PROBE_CHAINS];
#pragma unroll
  for (int k = 0; k < PROBE_CHAINS; ++k) a[k] = 1.0 + 1e-9 * (threadIdx.x + k);
#pragma unroll 4
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < PROBE_CHAINS; ++k) a[k] = __fma_rn(a[k], m, 1e-12);
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < PROBE_CHAINS; ++k) s += a[k];
  if (s == 12345.678) *sink = s;  // keeps the chains alive; never true
}

// ---- helpers shared with the grid schedule (c2b_fused.cuh) ----------------------------------------
// (u, v) of an observation, recomputed exactly as the cull kernel computed it
__device__ __forceinline__ double2 observe(const double *cam, double x, double y, double z) {
  V3 pc = project_world(cam, V3{x, y, z});
  double u, v;
  project(cam[12], cam[13], cam[14], pc, u, v);
  return make_double2(u, v);
}

// fallback for cameras that see more points than the shared-memory sort holds (SW_BLOCK_MAX): the scattered 64-bit keys were radix-sorted
// globally; split each key and recompute (u, v)
__global__ void k_write_sorted(const uint64_t *keys, uint64_t n, int pbits, const double *__restrict__ cams,
                               const double *__restrict__ p_aos, uint32_t *__restrict__ out_idx,
                               double2 *__restrict__ out_uv) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t key = keys[i];
  const uint64_t cam = key >> pbits, pt = key & ((1ull << pbits) - 1ull);
  double c[15];
#pragma unroll
  for (int k = 0; k < 15; ++k) c[k] = __ldg(&cams[15 * cam + k]);
  out_uv[i] = observe(c, p_aos[3 * pt], p_aos[3 * pt + 1], p_aos[3 * pt + 2]);
  out_idx[i] = (uint32_t)pt;
}

// ---- analytic occlusion of `city2ba synthetic` ------------------------------------------------------
// line_intersection 0.4.0: LineInterval::line_segment(a).relate(&line_segment(b))
// .unique_intersection(): parallel -> none; both parameters in [0,1] -> a.start + t*da.
__device__ __forceinline__ bool seg_intersect(double ax, double ay, double bx, double by, double cx,
                                              double cy, double dx, double dy, double &px,
                                              double &py) {
  double dax = dsub(bx, ax), day = dsub(by, ay), dbx = dsub(dx, cx), dby = dsub(dy, cy);
  double denom = dsub(dmul(dax, dby), dmul(day, dbx));
  if (denom == 0.0) return false;
  double sx = dsub(cx, ax), sy = dsub(cy, ay);
  // t = cross(q-p, s / rxs), u = cross(q-p, r / rxs): the crate divides the direction first
  double t = dsub(dmul(sx, ddiv(dby, denom)), dmul(sy, ddiv(dbx, denom)));
  double u = dsub(dmul(sx, ddiv(day, denom)), dmul(sy, ddiv(dax, denom)));
  if (t < 0.0 || t > 1.0 || u < 0.0 || u > 1.0) return false;
  px = dadd(ax, dmul(t, dax));
  py = dadd(ay, dmul(t, day));
  return true;
}

// src/synthetic.rs:52-98 including the un-squared term at :93 (NaN > 1e-8 is false)
__device__ __forceinline__ bool hits_in_block(double sx, double sy, double ex, double ey, long long bx,
                                              long long by, double L, double inset) {
  double be = dsub(L, inset), ox = dmul((double)bx, L), oy = dmul((double)by, L);
  double x0 = dadd(ox, inset), x1 = dadd(ox, be), y0 = dadd(oy, inset), y1 = dadd(oy, be);
  const double sides[4][4] = {{x0, y0, x0, y1}, {x0, y0, x1, y0}, {x1, y0, x1, y1}, {x0, y1, x1, y1}};
  bool any = false;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    double px, py;
    if (seg_intersect(sx, sy, ex, ey, sides[k][0], sides[k][1], sides[k][2], sides[k][3], px, py)) {
      double dx = dsub(ex, px);
      double v = dsqrt(dadd(dmul(dx, dx), dsub(ey, py)));
      if (v > 1e-8) any = true;
    }
  }
  return any;
}

// src/synthetic.rs:100-124
__device__ __forceinline__ bool hits_building(V3 c, V3 p, double L, double inset) {
  double sx = c.x, sy = c.z, ex = p.x, ey = p.z;
  long long cbx = (long long)trunc(ddiv(sx, L)), cby = (long long)trunc(ddiv(sy, L));
  long long pbx = (long long)trunc(ddiv(ex, L)), pby = (long long)trunc(ddiv(ey, L));
  long long x0 = cbx < pbx ? cbx : pbx, x1 = cbx < pbx ? pbx : cbx;
  long long y0 = cby < pby ? cby : pby, y1 = cby < pby ? pby : cby;
  for (long long x = x0; x <= x1; ++x)
    for (long long y = y0; y <= y1; ++y)
      if (hits_in_block(sx, sy, ex, ey, x, y, L, inset)) return true;
  return false;
}

__global__ void __launch_bounds__(256)
    k_analytic_occlusion(const uint64_t *__restrict__ keys, uint64_t n_cand, int pbits,
                         const double *__restrict__ cen_x, const double *__restrict__ cen_y,
                         const double *__restrict__ cen_z, const double *__restrict__ px,
                         const double *__restrict__ py, const double *__restrict__ pz, double L,
                         double inset, uint32_t *__restrict__ words) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if ((i & ~31ull) >= n_cand) return;
  bool vis = false;
  const uint64_t key = i < n_cand ? keys[i] : ~0ull;
  if (key != ~0ull) {
    const uint64_t cam = key >> pbits, pt = key & ((1ull << pbits) - 1ull);
    V3 c{cen_x[cam], cen_y[cam], cen_z[cam]};
    V3 p{px[pt], py[pt], pz[pt]};
    vis = !hits_building(c, p, L, inset);
  }
  const unsigned m = __ballot_sync(0xffffffffu, vis);
  if ((threadIdx.x & 31) == 0) words[i >> 5] = m;
}

}  // namespace c2b
