// c2b_sample.cuh — world-point generation on the GPU (generate_world_points_uniform,
// src/generate.rs:356-420): area-weighted triangle choice, a uniform point in the triangle, kept if
// some camera centre lies within max_dist.
//
// The reference runs a sequential rejection loop on thread_rng() (unseedable).  Here candidate k is a
// pure function of (seed, k): Philox4x32-10 counter (k, stream ST_SAMPLE, slot 0/1) gives the
// triangle variate and (rx, ry); the accepted candidates are kept in candidate order, so the output
// is deterministic, reproducible by any implementation of the same stream (the tests' CPU checker),
// and distributed like the reference's.  The
// cumulative-area table is a sequential f64 running sum built on the host (a fixed summation order),
// searched per candidate; camera proximity uses a uniform grid over the camera centres
// (cell = max_dist, 27 cells per query) with the reference's predicate dist^2 <= max_dist^2.
#pragma once
#include "c2b_common.cuh"
#include "c2b_math.cuh"
#include "c2b_noise.cuh"

namespace c2b {

enum { ST_SAMPLE = 16 };

struct CamGrid {
  double lo[3];
  double inv_h;
  int n[3];
};

__host__ __device__ __forceinline__ int cam_grid_coord(const CamGrid &g, int k, double x) {
  double t = floor((x - g.lo[k]) * g.inv_h);
  int c = t < 0.0 ? 0 : (t >= (double)g.n[k] ? g.n[k] - 1 : (int)t);
  if (!(t == t)) c = 0;
  return c;
}

struct SampleArgs {
  const float4 *tri_v;      // 3 float4 per triangle, ORIGINAL triangle order (w unused)
  const double *cdf;        // inclusive running sum of the areas, [nt]
  uint64_t nt;
  double total;
  CamGrid g;
  const uint32_t *cell_start;  // [cells + 1]
  const double *cen;           // camera centres sorted by cell, xyz interleaved
  double max_d2;
  uint64_t seed;
  uint64_t first;  // global index of candidate 0 of this round
  uint64_t n;      // candidates in this round
  double *cand;    // [3n] candidate points
  uint32_t *keep;  // [n] 1 = accepted
};

__device__ __forceinline__ double u01(uint32_t lo, uint32_t hi) {
  const uint64_t x = (uint64_t)lo | ((uint64_t)hi << 32);
  return dmul((double)(x >> 11), 1.1102230246251565e-16);
}

__global__ void __launch_bounds__(256) k_sample_points(SampleArgs a) {
  const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= a.n) return;
  const uint64_t k = a.first + j;
  uint32_t o[4], q[4];
  philox4x32_10((uint32_t)k, (uint32_t)(k >> 32), ST_SAMPLE, 0, (uint32_t)a.seed, (uint32_t)(a.seed >> 32), o);
  philox4x32_10((uint32_t)k, (uint32_t)(k >> 32), ST_SAMPLE, 1, (uint32_t)a.seed, (uint32_t)(a.seed >> 32), q);
  // WeightedIndex: first triangle whose running sum exceeds u * total
  const double target = dmul(u01(o[0], o[1]), a.total);
  uint64_t lo = 0, hi = a.nt - 1;
  while (lo < hi) {
    const uint64_t mid = (lo + hi) >> 1;
    if (a.cdf[mid] > target) hi = mid; else lo = mid + 1;
  }
  const float4 f0 = a.tri_v[3 * lo], f1 = a.tri_v[3 * lo + 1], f2 = a.tri_v[3 * lo + 2];
  // random_point_in_triangle (src/generate.rs:314-326)
  double rx = u01(q[0], q[1]), ry = u01(q[2], q[3]);
  if (dadd(rx, ry) > 1.0) {
    rx = dsub(1.0, rx);
    ry = dsub(1.0, ry);
  }
  const V3 v0{(double)f0.x, (double)f0.y, (double)f0.z};
  const V3 e1{dsub((double)f1.x, v0.x), dsub((double)f1.y, v0.y), dsub((double)f1.z, v0.z)};
  const V3 e2{dsub((double)f2.x, v0.x), dsub((double)f2.y, v0.y), dsub((double)f2.z, v0.z)};
  const V3 p{dadd(dadd(v0.x, dmul(rx, e1.x)), dmul(ry, e2.x)), dadd(dadd(v0.y, dmul(rx, e1.y)), dmul(ry, e2.y)),
             dadd(dadd(v0.z, dmul(rx, e1.z)), dmul(ry, e2.z))};
  a.cand[3 * j] = p.x;
  a.cand[3 * j + 1] = p.y;
  a.cand[3 * j + 2] = p.z;
  // any camera centre within max_dist?  cells of side >= max_dist: the 27 neighbours suffice
  bool near = false;
  if (p.x == p.x && p.y == p.y && p.z == p.z) {
    const int cx = cam_grid_coord(a.g, 0, p.x), cy = cam_grid_coord(a.g, 1, p.y), cz = cam_grid_coord(a.g, 2, p.z);
    for (int z = max(cz - 1, 0); z <= min(cz + 1, a.g.n[2] - 1) && !near; ++z)
      for (int y = max(cy - 1, 0); y <= min(cy + 1, a.g.n[1] - 1) && !near; ++y) {
        const uint32_t row = ((uint32_t)z * a.g.n[1] + y) * a.g.n[0];
        const uint32_t s = a.cell_start[row + max(cx - 1, 0)], e = a.cell_start[row + min(cx + 1, a.g.n[0] - 1) + 1];
        for (uint32_t c = s; c < e; ++c) {
          const V3 d{dsub(a.cen[3 * c], p.x), dsub(a.cen[3 * c + 1], p.y), dsub(a.cen[3 * c + 2], p.z)};
          if (mag2(d) <= a.max_d2) {
            near = true;
            break;
          }
        }
      }
  }
  a.keep[j] = near ? 1u : 0u;
}

// accepted candidates, in candidate order, appended at out[base...] up to `limit` points in total;
// rank = exclusive scan of keep
__global__ void k_sample_compact(const double *__restrict__ cand, const uint32_t *__restrict__ keep,
                                 const uint32_t *__restrict__ rank, uint64_t n, uint64_t base, uint64_t limit,
                                 double *__restrict__ out, unsigned long long *__restrict__ last_used) {
  const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n || !keep[j]) return;
  const uint64_t pos = base + rank[j];
  if (pos >= limit) return;
  out[3 * pos] = cand[3 * j];
  out[3 * pos + 1] = cand[3 * j + 1];
  out[3 * pos + 2] = cand[3 * j + 2];
  if (pos == limit - 1) *last_used = j;  // the candidate that completed the request
}

}  // namespace c2b
