// c2b_traverse.cuh — kernel (2): occlusion rays against the LBVH (replaces Embree's
// rtcOccluded1M, src/generate.rs:472).
//
// Warp-cooperative, stackless: a warp owns 32 consecutive candidates (rays that share or nearly
// share a camera origin) and walks the threaded pre-order tree ONCE for all of them.  The node
// index is warp-uniform, so every node / triangle fetch is a single broadcast load and there is
// no divergence in the walk: a node is entered if ANY live lane's ray overlaps its box
// (__ballot_sync), otherwise the warp jumps to the node's escape index.  Lanes whose ray has been
// occluded drop out of the vote; the walk ends when none is left or the escape index runs off
// the tree.  Any-hit is an OR over triangles, and the box test is conservative with respect to
// the watertight triangle test, so the result is independent of the tree's shape.
#pragma once
#include "c2b_common.cuh"
#include "c2b_math.cuh"

namespace c2b {

struct BoxRay {
  float ix, iy, iz;     // 1/dir, magnitude clamped to 1e30 so that o*i stays finite
  float cx, cy, cz;     // -org * inv
  float px, py, pz;     // slab padding in t units: pad * |inv|
  float tfar;
};

// Per-ray set-up of the conservative slab test.  `scene_absmax` = largest |coordinate| of the
// scene bounds.  Every box is (implicitly) inflated by pad = 4e-6 * (|org|_inf + scene_absmax)
// on each side, which covers (a) the rounding of fma(lo, inv, -org*inv) — at most
// 6e-8 * |org*inv| in t — and (b) the few-ulp geometric slop of the watertight triangle test,
// which works on vertices translated by org (magnitudes <= |org| + scene_absmax).  The residual
// relative error of t (one rounding) is absorbed by the 1e-5 slack of the final compare.
__device__ __forceinline__ BoxRay make_box_ray(const Ray &r, float scene_absmax) {
  BoxRay b;
  const float big = 1e30f;
  float ix = 1.0f / r.dx, iy = 1.0f / r.dy, iz = 1.0f / r.dz;
  b.ix = fabsf(ix) > big ? copysignf(big, r.dx) : ix;
  b.iy = fabsf(iy) > big ? copysignf(big, r.dy) : iy;
  b.iz = fabsf(iz) > big ? copysignf(big, r.dz) : iz;
  b.cx = -r.ox * b.ix;
  b.cy = -r.oy * b.iy;
  b.cz = -r.oz * b.iz;
  const float pad = 4e-6f * (fmaxf(fabsf(r.ox), fmaxf(fabsf(r.oy), fabsf(r.oz))) + scene_absmax) + 1e-30f;
  b.px = pad * fabsf(b.ix);
  b.py = pad * fabsf(b.iy);
  b.pz = pad * fabsf(b.iz);
  b.tfar = r.tfar;
  return b;
}

// 2 FMA + 4 FMNMX + 2 FADD per axis.  fminf/fmaxf return the non-NaN operand, so an axis whose
// products overflow to inf - inf is simply ignored (conservative).
__device__ __forceinline__ bool box_overlap(const BoxRay &r, float4 lo, float4 hi) {
  float tmin = 0.0f, tmax = r.tfar;
  {
    float t0 = fmaf(lo.x, r.ix, r.cx), t1 = fmaf(hi.x, r.ix, r.cx);
    tmin = fmaxf(tmin, fminf(t0, t1) - r.px);
    tmax = fminf(tmax, fmaxf(t0, t1) + r.px);
  }
  {
    float t0 = fmaf(lo.y, r.iy, r.cy), t1 = fmaf(hi.y, r.iy, r.cy);
    tmin = fmaxf(tmin, fminf(t0, t1) - r.py);
    tmax = fminf(tmax, fmaxf(t0, t1) + r.py);
  }
  {
    float t0 = fmaf(lo.z, r.iz, r.cz), t1 = fmaf(hi.z, r.iz, r.cz);
    tmin = fmaxf(tmin, fminf(t0, t1) - r.pz);
    tmax = fminf(tmax, fmaxf(t0, t1) + r.pz);
  }
  return tmin <= tmax * 1.00001f + 1e-30f;
}

// warp-cooperative any-hit.  `alive` = this lane has a ray that is not yet occluded.
// Returns true if this lane's ray is occluded.
// MT = the Moeller-Trumbore predicate (c2b_math.cuh: ray_triangle_mt) instead of the watertight one
template <bool COUNT, bool MT = false>
__device__ __forceinline__ bool warp_any_hit(const float4 *__restrict__ nodes,
                                             const float4 *__restrict__ tris, int n_nodes,
                                             const Ray &ray, bool alive, float scene_absmax,
                                             unsigned long long *counters) {
  const BoxRay br = make_box_ray(ray, scene_absmax);
  // NaN / negative tfar can never satisfy 0 < t <= tfar; a NaN direction can never hit either
  alive = alive && (ray.tfar >= 0.0f) && (ray.dx == ray.dx) && (ray.dy == ray.dy) && (ray.dz == ray.dz);
  bool occluded = false;
  int node = 0;
  unsigned n_vis = 0, n_tri = 0;
  while (node < n_nodes) {
    if (__ballot_sync(0xffffffffu, alive) == 0u) break;
    const float4 lo = __ldg(&nodes[2 * node]);
    const float4 hi = __ldg(&nodes[2 * node + 1]);
    if (COUNT) ++n_vis;
    const bool hit = alive && box_overlap(br, lo, hi);
    const unsigned hm = __ballot_sync(0xffffffffu, hit);
    const int esc = __float_as_int(lo.w);
    if (hm == 0u) {
      node = esc;
      continue;
    }
    const int leaf = __float_as_int(hi.w);
    if (leaf >= 0) {
      const float4 v0 = __ldg(&tris[3 * leaf]);
      const float4 v1 = __ldg(&tris[3 * leaf + 1]);
      const float4 v2 = __ldg(&tris[3 * leaf + 2]);
      if (COUNT) ++n_tri;
      if (hit && (MT ? ray_triangle_mt(ray, v0.x, v0.y, v0.z, v1.x, v1.y, v1.z, v2.x, v2.y, v2.z)
                     : ray_triangle(ray, v0.x, v0.y, v0.z, v1.x, v1.y, v1.z, v2.x, v2.y, v2.z))) {
        occluded = true;
        alive = false;
      }
      node = esc;  // == node + 1 for a leaf
    } else {
      node = node + 1;
    }
  }
  if (COUNT && (threadIdx.x & 31) == 0) {
    atomicAdd(&counters[2], (unsigned long long)n_vis);
    atomicAdd(&counters[3], (unsigned long long)n_tri);
  }
  return occluded;
}

struct TraverseArgs {
  const float4 *nodes;
  const float4 *tris;
  int n_nodes;
  const uint64_t *keys;  // candidate keys (sorted, or the pool with POOL_SENTINEL padding)
  uint64_t n_cand;
  float scene_absmax;
  int pbits;
  const double *cen_x, *cen_y, *cen_z;
  const double *p_aos;  // xyz records, original point order
  int endpoint_guard_rel;
  uint32_t *vis_words;  // bit i of word w set  <=>  candidate 32*w+i is VISIBLE
  unsigned long long *counters;
};

template <bool COUNT, bool MT = false>
__global__ void __launch_bounds__(256) k_traverse(TraverseArgs a) {
  const int lane = threadIdx.x & 31;
  const uint64_t w = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const uint64_t i = w * 32 + lane;
  if (w * 32 >= a.n_cand) return;
  bool have = i < a.n_cand;
  Ray ray;
  ray.ox = ray.oy = ray.oz = 0.0f;
  ray.dx = ray.dy = ray.dz = 1.0f;
  ray.tfar = -1.0f;
  const uint64_t key = have ? a.keys[i] : ~0ull;
  have = have && key != ~0ull;  // padding slot of an aligned chunk
  if (have) {
    const uint64_t cam = key >> a.pbits, pt = key & ((1ull << a.pbits) - 1ull);
    V3 c{a.cen_x[cam], a.cen_y[cam], a.cen_z[cam]};
    V3 p{a.p_aos[3 * pt], a.p_aos[3 * pt + 1], a.p_aos[3 * pt + 2]};
    ray = make_ray(c, p, a.endpoint_guard_rel != 0);
  }
  const bool occ = warp_any_hit<COUNT, MT>(a.nodes, a.tris, a.n_nodes, ray, have, a.scene_absmax, a.counters);
  const unsigned vm = __ballot_sync(0xffffffffu, have && !occ);
  if (lane == 0) a.vis_words[w] = vm;
}

// ---- grid schedule with the Moeller-Trumbore predicate ---------------------------------------------------
// The fused kernel's packet machinery (origin-relative records, direction-box prefilter) is built around the
// watertight edge functions.  For the non-default MT predicate the fused pass runs WITHOUT occlusion, which
// leaves every camera's candidates in its scratch slice, and this kernel — one warp per slot, 32 candidates at
// a time through the generic walk — keeps the visible ones, compacted in place (the write position never
// passes the read position, and a chunk is read completely before anything is written).
struct FilterArgs {
  const float4 *nodes;
  const float4 *tris;
  int n_nodes;
  float scene_absmax;
  const double *cen_x, *cen_y, *cen_z;
  const double *p_aos;
  uint64_t slots;
  int parts_log2;
  int endpoint_guard_rel;
  const uint32_t *ev_off;  // [slots + 1] scratch slice starts
  uint32_t *scratch_idx;
  uint32_t *vis_count;     // [slots] candidates in, visible out
  unsigned long long *counters;
};

__global__ void __launch_bounds__(256) k_filter_candidates_mt(FilterArgs a) {
  const int lane = threadIdx.x & 31;
  const uint64_t slot = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (slot >= a.slots) return;
  const uint64_t cam = slot >> a.parts_log2;
  const uint32_t n = a.vis_count[slot];
  uint32_t *list = a.scratch_idx + a.ev_off[slot];
  const V3 c{a.cen_x[cam], a.cen_y[cam], a.cen_z[cam]};
  uint32_t kept = 0;
  for (uint32_t base = 0; base < n; base += 32) {
    const bool have = base + lane < n;
    uint32_t pt = 0;
    Ray ray;
    ray.ox = ray.oy = ray.oz = 0.0f;
    ray.dx = ray.dy = ray.dz = 1.0f;
    ray.tfar = -1.0f;
    if (have) {
      pt = list[base + lane];
      ray = make_ray(c, V3{a.p_aos[3 * (uint64_t)pt], a.p_aos[3 * (uint64_t)pt + 1], a.p_aos[3 * (uint64_t)pt + 2]},
                     a.endpoint_guard_rel != 0);
    }
    const bool occ = warp_any_hit<false, true>(a.nodes, a.tris, a.n_nodes, ray, have, a.scene_absmax, a.counters);
    const bool vis = have && !occ;
    const unsigned m = __ballot_sync(0xffffffffu, vis);
    __syncwarp();
    if (vis) list[kept + __popc(m & ((1u << lane) - 1u))] = pt;
    kept += __popc(m);
    __syncwarp();
  }
  if (lane == 0) a.vis_count[slot] = kept;
}

// ---- per-camera triangle lists ------------------------------------------------------------------------
// All rays of a camera start at its centre and are shorter than max_dist, so every triangle any of
// them can hit lies in the cube [c - R, c + R]^3.  k_cam_trilist walks the tree once per camera
// (one warp each, lane = node) and records the leaves whose box overlaps that cube, up to `cap`
// per camera (more => TRILIST_OVERFLOW, the camera's packets use the generic walk).  A packet then
// needs no tree walk at all (packet_any_hit, c2b_fused.cuh): lanes test 32 list entries at a time
// against the packet's bounding box (lane = triangle), and only the surviving triangles are run
// through the watertight test (lane = ray).  The result is identical: a triangle that a ray hits
// overlaps both the cube and the packet box (both inflated by the same conservative pad as the
// slab test).
constexpr uint32_t TRILIST_OVERFLOW = 0xffffffffu;

__device__ __forceinline__ float conservative_pad(float ox, float oy, float oz, float scene_absmax) {
  return 4e-6f * (fmaxf(fabsf(ox), fmaxf(fabsf(oy), fabsf(oz))) + scene_absmax) + 1e-30f;
}

// serial stackless walk (one thread): the fall-back of the warp version below
__device__ __forceinline__ uint32_t trilist_serial(const float4 *__restrict__ nodes, int n_nodes, float lx, float ly,
                                                   float lz, float hx, float hy, float hz, uint32_t cap,
                                                   uint32_t *__restrict__ mylist) {
  uint32_t n = 0;
  int node = 0;
  while (node < n_nodes) {
    const float4 lo = __ldg(&nodes[2 * node]);
    const float4 hi = __ldg(&nodes[2 * node + 1]);
    // written so that NaN compares keep the node (conservative)
    const bool outside = lo.x > hx || hi.x < lx || lo.y > hy || hi.y < ly || lo.z > hz || hi.z < lz;
    if (outside) {
      node = __float_as_int(lo.w);
    } else if (__float_as_int(hi.w) >= 0) {
      if (n < cap) mylist[n] = (uint32_t)node;
      ++n;
      if (n > cap) break;
      node = __float_as_int(lo.w);
    } else {
      node = node + 1;
    }
  }
  return n;
}

// one thread per camera: the cheapest form when there are enough cameras to fill the GPU (cfg4, one GPU:
// 0.075 ms for 99,840 cameras against 0.15 ms for the warp form below)
__global__ void __launch_bounds__(128)
    k_cam_trilist(const float4 *__restrict__ nodes, int n_nodes, const double *__restrict__ cen_x,
                  const double *__restrict__ cen_y, const double *__restrict__ cen_z, uint64_t C,
                  float rmax, float scene_absmax, uint32_t cap, uint32_t *__restrict__ list,
                  uint32_t *__restrict__ count, unsigned long long *__restrict__ n_overflow) {
  const uint64_t cam = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (cam >= C) return;
  const float ox = __double2float_rn(cen_x[cam]), oy = __double2float_rn(cen_y[cam]),
              oz = __double2float_rn(cen_z[cam]);
  const float r = rmax + conservative_pad(ox, oy, oz, scene_absmax);
  const uint32_t n = trilist_serial(nodes, n_nodes, ox - r, oy - r, oz - r, ox + r, oy + r, oz + r, cap, list + cam * cap);
  count[cam] = n <= cap ? n : TRILIST_OVERFLOW;
  if (n > cap) atomicAdd(n_overflow, 1ull);
}

// One warp per camera, lane = node, for calls with few cameras (multi-GPU shards, early batches), where
// the serial walk is a chain of dependent loads that takes the same 0.07 ms for 12,480 cameras as for
// 99,840: the warp keeps a LIFO of node indices in shared memory, pops up to 32 per step, every lane
// tests one node's box against the camera's cube, overlapping leaves are appended to the list (ballot
// order), overlapping inner nodes push both children (left = node + 1, right child index in the node's
// second w).
constexpr int TL_WARPS = 4;
constexpr int TL_LIFO = 320;

__global__ void __launch_bounds__(TL_WARPS * 32)
    k_cam_trilist_warp(const float4 *__restrict__ nodes, int n_nodes, const double *__restrict__ cen_x,
                  const double *__restrict__ cen_y, const double *__restrict__ cen_z, uint64_t C,
                  float rmax, float scene_absmax, uint32_t cap, uint32_t *__restrict__ list,
                  uint32_t *__restrict__ count, unsigned long long *__restrict__ n_overflow) {
  __shared__ uint32_t s_lifo[TL_WARPS][TL_LIFO];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t cam = (uint64_t)blockIdx.x * TL_WARPS + warp;
  if (cam >= C) return;
  const float ox = __double2float_rn(cen_x[cam]), oy = __double2float_rn(cen_y[cam]),
              oz = __double2float_rn(cen_z[cam]);
  const float r = rmax + conservative_pad(ox, oy, oz, scene_absmax);
  const float lx = ox - r, ly = oy - r, lz = oz - r, hx = ox + r, hy = oy + r, hz = oz + r;
  uint32_t *mylist = list + cam * cap;
  uint32_t *lifo = s_lifo[warp];
  const unsigned lt = (1u << lane) - 1u;
  uint32_t n = 0;
  int sp = n_nodes > 0 ? 1 : 0;
  bool serial = false;
  if (lane == 0) lifo[0] = 0u;
  __syncwarp();
  while (sp > 0) {
    const int take = sp < 32 ? sp : 32;
    bool leaf = false, inner = false;
    uint32_t node = 0, right = 0;
    if (lane < take) {
      node = lifo[sp - take + lane];
      const float4 lo = __ldg(&nodes[2 * node]);
      const float4 hi = __ldg(&nodes[2 * node + 1]);
      // written so that NaN compares keep the node (conservative)
      const bool outside = lo.x > hx || hi.x < lx || lo.y > hy || hi.y < ly || lo.z > hz || hi.z < lz;
      if (!outside) {
        const int slot = __float_as_int(hi.w);
        leaf = slot >= 0;
        inner = !leaf;
        right = (uint32_t)(-slot - 1);
      }
    }
    __syncwarp();  // every lane has read its node before anyone pushes into the same slots
    sp -= take;
    const unsigned lm = __ballot_sync(0xffffffffu, leaf), im = __ballot_sync(0xffffffffu, inner);
    if (leaf) {
      const uint32_t pos = n + __popc(lm & lt);
      if (pos < cap) mylist[pos] = node;
    }
    n += __popc(lm);
    if (n > cap) break;
    if (sp + 2 * __popc(im) > TL_LIFO) {
      serial = true;
      break;
    }
    if (inner) {
      const int pos = sp + 2 * __popc(im & lt);
      lifo[pos] = node + 1u;
      lifo[pos + 1] = right;
    }
    sp += 2 * __popc(im);
    __syncwarp();
  }
  if (serial) {  // the LIFO would overflow: one lane walks the tree without memory
    if (lane == 0) n = trilist_serial(nodes, n_nodes, lx, ly, lz, hx, hy, hz, cap, mylist);
    n = __shfl_sync(0xffffffffu, n, 0);
  }
  if (lane == 0) {
    count[cam] = n <= cap ? n : TRILIST_OVERFLOW;
    if (n > cap) atomicAdd(n_overflow, 1ull);
  }
}

// ---- Embree-shaped ray batch (parity tooling): AoS 48-byte rays, tfar = -inf on hit ---------------
template <bool COUNT>
__global__ void __launch_bounds__(256) k_occluded_rays(const float4 *__restrict__ nodes,
                                                       const float4 *__restrict__ tris, int n_nodes,
                                                       c2b_ray48 *rays, uint64_t n, float scene_absmax,
                                                       unsigned long long *counters) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if ((i & ~31ull) >= n) return;
  const bool have = i < n;
  Ray ray;
  ray.ox = ray.oy = ray.oz = 0.0f;
  ray.dx = ray.dy = ray.dz = 1.0f;
  ray.tfar = -1.0f;
  if (have) {
    ray.ox = rays[i].org_x;
    ray.oy = rays[i].org_y;
    ray.oz = rays[i].org_z;
    ray.dx = rays[i].dir_x;
    ray.dy = rays[i].dir_y;
    ray.dz = rays[i].dir_z;
    ray.tfar = rays[i].tfar;
  }
  const bool occ = warp_any_hit<COUNT>(nodes, tris, n_nodes, ray, have, scene_absmax, counters);
  if (have && occ) rays[i].tfar = -INFINITY;
}

// ---- closest hit (Embree rtcIntersect1, src/generate.rs:253-262: camera placement rays) -----------
// Same warp-cooperative stackless walk as warp_any_hit, but a lane keeps walking after a hit and
// remembers the smallest t.  Every triangle is tested against the ray's ORIGINAL tfar and the
// result is min over the hits of T/det, so it equals the brute-force minimum over all triangles bit
// for bit; the running best only tightens the (conservative) box test.
__device__ __forceinline__ float warp_closest_hit(const float4 *__restrict__ nodes,
                                                  const float4 *__restrict__ tris, int n_nodes,
                                                  const Ray &ray, bool alive, float scene_absmax) {
  BoxRay br = make_box_ray(ray, scene_absmax);
  alive = alive && (ray.tfar >= 0.0f) && (ray.dx == ray.dx) && (ray.dy == ray.dy) && (ray.dz == ray.dz);
  float best = INFINITY;
  int node = 0;
  while (node < n_nodes) {
    const float4 lo = __ldg(&nodes[2 * node]);
    const float4 hi = __ldg(&nodes[2 * node + 1]);
    const bool hit = alive && box_overlap(br, lo, hi);
    const unsigned hm = __ballot_sync(0xffffffffu, hit);
    const int esc = __float_as_int(lo.w);
    if (hm == 0u) {
      node = esc;
      continue;
    }
    const int leaf = __float_as_int(hi.w);
    if (leaf >= 0) {
      const float4 v0 = __ldg(&tris[3 * leaf]);
      const float4 v1 = __ldg(&tris[3 * leaf + 1]);
      const float4 v2 = __ldg(&tris[3 * leaf + 2]);
      float t;
      if (hit && ray_triangle(ray, v0.x, v0.y, v0.z, v1.x, v1.y, v1.z, v2.x, v2.y, v2.z, &t) && t < best) {
        best = t;
        br.tfar = fminf(ray.tfar, best);
      }
      node = esc;  // == node + 1 for a leaf
    } else {
      node = node + 1;
    }
  }
  return best;
}

// Embree-shaped closest-hit batch: on a hit tfar = t and flags = 1, otherwise flags = 0
__global__ void __launch_bounds__(256) k_intersect_rays(const float4 *__restrict__ nodes,
                                                        const float4 *__restrict__ tris, int n_nodes,
                                                        c2b_ray48 *rays, uint64_t n, float scene_absmax) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if ((i & ~31ull) >= n) return;
  const bool have = i < n;
  Ray ray;
  ray.ox = ray.oy = ray.oz = 0.0f;
  ray.dx = ray.dy = ray.dz = 1.0f;
  ray.tfar = -1.0f;
  if (have) {
    ray.ox = rays[i].org_x;
    ray.oy = rays[i].org_y;
    ray.oz = rays[i].org_z;
    ray.dx = rays[i].dir_x;
    ray.dy = rays[i].dir_y;
    ray.dz = rays[i].dir_z;
    ray.tfar = rays[i].tfar;
  }
  const float t = warp_closest_hit(nodes, tris, n_nodes, ray, have, scene_absmax);
  if (have) {
    const bool hit = t < INFINITY;
    if (hit) rays[i].tfar = t;
    rays[i].flags = hit ? 1u : 0u;
  }
}

}  // namespace c2b
