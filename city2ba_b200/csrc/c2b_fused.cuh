// c2b_fused.cuh — the grid schedule of c2b_visibility_graph as ONE pass per camera:
// cull + projection (src/generate.rs:446-454), ray construction (:456-464), occlusion (:472) and
// the visible-point list (:473-478) without a candidate pool in between.
//
//   k_cam_plan          one thread per camera: how many grid-ordered points its rows hold (an upper
//                       bound of its visible count) -> exclusive scan -> the camera's slice of the
//                       scratch index array.
//   k_visibility_fused  one warp per camera.  The warp scans the x-contiguous cell rows its
//                       max_dist ball touches (each row trimmed to the half-space in front of the
//                       camera), evaluates the exact f64 predicate per lane and stages the
//                       survivors' grid positions in shared memory.  Every 32 survivors form a ray
//                       packet that is resolved at once: the rays are built from the coordinates
//                       just read (L1 hits), the camera's leaf list (k_cam_trilist) is filtered
//                       against the packet's bounding box with lane = triangle, the survivors'
//                       origin-relative triangle records (TriRec, 3 x float4) are computed ONCE for
//                       the packet and parked in shared memory, and then lane = ray runs the 9-FMA
//                       edge-function test per record.  Visible point indices go to the camera's
//                       scratch slice; the camera's visible count needs no atomics.
//   k_sort_write        one warp per camera: register bitonic sort of the visible indices
//                       (ascending point index, src/generate.rs:446), (u, v) recomputed with the
//                       cull kernel's device function (bit-identical), CSR records written coalesced.
#pragma once
#include "c2b_common.cuh"
#include "c2b_compact.cuh"
#include "c2b_cull.cuh"
#include "c2b_math.cuh"
#include "c2b_traverse.cuh"

namespace c2b {

struct FusedArgs {
  // cameras
  const double *cams;                   // [15*C]
  const double *cen_x, *cen_y, *cen_z;  // [C]
  uint64_t C;
  // grid-ordered points
  const double *gx, *gy, *gz;  // [P]
  const uint32_t *gidx;        // [P] original index
  const uint32_t *cell_start;  // [n_cells + 1]
  GridDesc g;
  double max_dist;
  double t_star;  // m2 < t_star  <=>  sqrt_rn(m2) < max_dist (exact)
  // occlusion
  const float4 *nodes;
  const float4 *tris;
  int n_nodes;
  float scene_absmax;
  const uint32_t *tri_list;   // [C * tri_cap] leaf node indices (k_cam_trilist)
  const uint32_t *tri_count;  // [C]
  uint32_t tri_cap;
  int endpoint_guard_rel;
  double block_length, block_inset;  // analytic occlusion (src/synthetic.rs:52-124)
  // plan + output
  uint32_t *ev_count;         // [C+1] points on the camera's rows (k_cam_plan), then its exclusive scan
  uint32_t *scratch_idx;      // visible point indices, camera c at [ev_off[c], ev_off[c] + vis_count[c])
  uint32_t *vis_count;        // [C+1]
  unsigned long long *counters;  // [1] pairs evaluated, [2] list entries / nodes, [3] warp triangle tests,
                                 // [4] candidates (rays), [5] scratch slice overflow flag
};

// camera_cell_range / row_span are __noinline__ on purpose: k_cam_plan sizes each camera's scratch
// slice from the very rows k_visibility_fused later scans, so both must run the SAME machine code
// (an FMA contracted in one inlined copy and not in the other could move a row end by one cell).

// cells whose points can lie within max_dist of the centre; false = none
__device__ __noinline__ bool camera_cell_range(const GridDesc &g, const double cc[3], double max_dist,
                                                  int lo[3], int hi[3]) {
  bool empty = !(max_dist > 0.0);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    // every point with |p - c| < max_dist has c_k - r - eps < p_k < c_k + r + eps, and grid_coord is
    // monotone, so [coord(a0), coord(a1)] holds its cell; eps covers the rounding of a0 / a1
    const double eps = 1e-9 * (fabs(cc[k]) + fabs(max_dist)) + 1e-300;
    const double a0 = cc[k] - max_dist - eps, a1 = cc[k] + max_dist + eps;
    if (a0 > g.max_c[k] || a1 < g.min_c[k]) empty = true;  // ball misses the populated slab
    lo[k] = grid_coord(g, k, a0);
    hi[k] = grid_coord(g, k, a1);
  }
  if (!(cc[0] == cc[0] && cc[1] == cc[1] && cc[2] == cc[2])) empty = true;  // NaN centre sees nothing
  return !empty;
}

// "in front of the camera" is pc.z = r2x*x + r2y*y + r2z*z + tz <= 0 (src/generate.rs:450): linear in
// x along a row of cells, so the row [x0, x1] is trimmed to the x-range that can hold such a point.
// false = nothing on this row can be in front.
__device__ __noinline__ bool row_span(const GridDesc &g, double r2x, double r2y, double r2z, double tz,
                                         double ccx, double max_dist, int y, int z, int &x0, int &x1) {
  const double cell_h = 1.0 / g.inv_h;
  // bounds of the row's cells in y and z; edge cells also hold the clamped coordinates, so they
  // extend to the data bounds
  const double sl = 1e-6 * cell_h;
  const double y0 = y == 0 ? g.min_c[1] : g.lo[1] + y * cell_h - sl;
  const double y1 = y == g.n[1] - 1 ? g.max_c[1] : g.lo[1] + (y + 1) * cell_h + sl;
  const double z0 = z == 0 ? g.min_c[2] : g.lo[2] + z * cell_h - sl;
  const double z1 = z == g.n[2] - 1 ? g.max_c[2] : g.lo[2] + (z + 1) * cell_h + sl;
  // smallest value r2y*y + r2z*z + tz can take on the row (0 * inf is avoided explicitly)
  const double my = r2y == 0.0 ? 0.0 : r2y * (r2y > 0.0 ? y0 : y1);
  const double mz = r2z == 0.0 ? 0.0 : r2z * (r2z > 0.0 ? z0 : z1);
  const double bmin = my + mz + tz;
  const double xr = fabs(r2x) * (fabs(ccx) + fabs(max_dist));
  const double mag = fabs(my) + fabs(mz) + fabs(tz) + xr;
  if (bmin == bmin && fabs(bmin) < INFINITY) {
    const double slack = 1e-9 * mag + 1e-300;
    if (xr <= slack) {
      if (bmin > 2.0 * slack) return false;  // the whole row is behind the camera
    } else {
      const double xlim = (-(bmin - slack)) / r2x;  // r2x*x <= -(bmin - slack)
      if (xlim == xlim) {
        if (r2x > 0.0) {
          const double xe = xlim + 1e-9 * (fabs(xlim) + cell_h);
          if (xe < g.lo[0]) return false;  // nothing in front on this row
          const int xc = grid_coord(g, 0, xe);
          x1 = xc < x1 ? xc : x1;
        } else {
          const double xe = xlim - 1e-9 * (fabs(xlim) + cell_h);
          if (xe > g.max_c[0]) return false;
          const int xc = grid_coord(g, 0, xe);
          x0 = xc > x0 ? xc : x0;
        }
        if (x0 > x1) return false;
      }
    }
  }
  return true;
}

// ---- plan -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_cam_plan(FusedArgs a) {
  const uint64_t cam = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t n = 0;
  if (cam < a.C) {
    const double cc[3] = {a.cen_x[cam], a.cen_y[cam], a.cen_z[cam]};
    int lo[3], hi[3];
    if (camera_cell_range(a.g, cc, a.max_dist, lo, hi)) {
      const double *c = a.cams + 15 * cam;
      const double r2x = c[2], r2y = c[5], r2z = c[8], tz = c[11];
      for (int z = lo[2]; z <= hi[2]; ++z)
        for (int y = lo[1]; y <= hi[1]; ++y) {
          int x0 = lo[0], x1 = hi[0];
          if (!row_span(a.g, r2x, r2y, r2z, tz, cc[0], a.max_dist, y, z, x0, x1)) continue;
          const uint32_t row = ((uint32_t)z * a.g.n[1] + y) * a.g.n[0];
          n += a.cell_start[row + x1 + 1] - a.cell_start[row + x0];
        }
    }
  }
  if (cam <= a.C) a.ev_count[cam] = (uint32_t)n;  // n <= P < 2^32; slot C = 0 closes the scan
  // 64-bit total, so the host can tell when the u32 scan would wrap
  unsigned long long s = n;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0 && s) atomicAdd(&a.counters[1], s);
}

// ---- the fused pass ---------------------------------------------------------------------------------
constexpr int FU_WARPS = 8;
constexpr int FU_STAGE = 64;

enum { FU_OCC_MESH = 0, FU_OCC_NONE = 1, FU_OCC_ANALYTIC = 2 };

// float <-> unsigned with the same ordering (no NaNs), for REDUX min / max
__device__ __forceinline__ unsigned f2ord(float f) {
  const unsigned b = __float_as_uint(f);
  return b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
  return __uint_as_float(u ^ ((u >> 31) ? 0x80000000u : 0xffffffffu));
}

// list-driven any-hit for a packet of rays that share the origin (ox, oy, oz).  rec = 32 x 3 float4
// of this warp.  Returns true when this lane's ray is occluded.
template <bool COUNT>
__device__ __forceinline__ bool packet_any_hit(const FusedArgs &a, const uint32_t *__restrict__ mylist,
                                               uint32_t n_list, const Ray &ray, bool have, float ox,
                                               float oy, float oz, float4 *rec, int lane,
                                               unsigned &n_vis, unsigned &n_tri) {
  bool alive = have && (ray.tfar >= 0.0f) && (ray.dx == ray.dx) && (ray.dy == ray.dy) && (ray.dz == ray.dz);
  bool occ = false;
  // bounding box of the packet's ray segments (origin + end points), conservatively padded
  float lx = INFINITY, ly = INFINITY, lz = INFINITY, hx = -INFINITY, hy = -INFINITY, hz = -INFINITY;
  if (alive) {
    const float pad = conservative_pad(ox, oy, oz, a.scene_absmax) + 4e-6f * ray.tfar;
    const float ex = fmaf(ray.dx, ray.tfar, ox), ey = fmaf(ray.dy, ray.tfar, oy), ez = fmaf(ray.dz, ray.tfar, oz);
    lx = fminf(ox, ex) - pad;
    ly = fminf(oy, ey) - pad;
    lz = fminf(oz, ez) - pad;
    hx = fmaxf(ox, ex) + pad;
    hy = fmaxf(oy, ey) + pad;
    hz = fmaxf(oz, ez) + pad;
  }
  if (__ballot_sync(0xffffffffu, alive) == 0u) return false;
  const float blx = ord2f(__reduce_min_sync(0xffffffffu, f2ord(lx)));
  const float bly = ord2f(__reduce_min_sync(0xffffffffu, f2ord(ly)));
  const float blz = ord2f(__reduce_min_sync(0xffffffffu, f2ord(lz)));
  const float bhx = ord2f(__reduce_max_sync(0xffffffffu, f2ord(hx)));
  const float bhy = ord2f(__reduce_max_sync(0xffffffffu, f2ord(hy)));
  const float bhz = ord2f(__reduce_max_sync(0xffffffffu, f2ord(hz)));
  for (uint32_t base = 0; base < n_list; base += 32) {
    const uint32_t j = base + lane;
    bool overlap = false;
    int slot = 0;
    if (j < n_list) {
      const uint32_t node = mylist[j];
      const float4 lo = __ldg(&a.nodes[2 * node]);
      const float4 hi = __ldg(&a.nodes[2 * node + 1]);
      slot = __float_as_int(hi.w);
      overlap = !(lo.x > bhx || hi.x < blx || lo.y > bhy || hi.y < bly || lo.z > bhz || hi.z < blz);
    }
    const unsigned m = __ballot_sync(0xffffffffu, overlap);
    if (COUNT) n_vis += min(32u, n_list - base);
    if (m == 0u) continue;
    if (overlap) {
      const float4 v0 = __ldg(&a.tris[3 * slot]);
      const float4 v1 = __ldg(&a.tris[3 * slot + 1]);
      const float4 v2 = __ldg(&a.tris[3 * slot + 2]);
      const TriRec t = tri_record(ox, oy, oz, v0.x, v0.y, v0.z, v1.x, v1.y, v1.z, v2.x, v2.y, v2.z);
      float4 *r = rec + 3 * __popc(m & ((1u << lane) - 1u));
      r[0] = make_float4(t.ux, t.uy, t.uz, t.vx);
      r[1] = make_float4(t.vy, t.vz, t.wx, t.wy);
      r[2] = make_float4(t.wz, t.T, 0.0f, 0.0f);
    }
    __syncwarp();
    const int cnt = __popc(m);
    for (int k = 0; k < cnt; ++k) {
      const float4 r0 = rec[3 * k], r1 = rec[3 * k + 1], r2 = rec[3 * k + 2];
      TriRec t;
      t.ux = r0.x; t.uy = r0.y; t.uz = r0.z;
      t.vx = r0.w; t.vy = r1.x; t.vz = r1.y;
      t.wx = r1.z; t.wy = r1.w; t.wz = r2.x;
      t.T = r2.y;
      if (COUNT) ++n_tri;
      if (alive && ray_tri_record(ray.dx, ray.dy, ray.dz, ray.tfar, t)) {
        occ = true;
        alive = false;
      }
      if (__ballot_sync(0xffffffffu, alive) == 0u) break;
    }
    __syncwarp();
    if (__ballot_sync(0xffffffffu, alive) == 0u) break;
  }
  return occ;
}

template <int OCC, bool COUNT>
__global__ void __launch_bounds__(FU_WARPS * 32) k_visibility_fused(FusedArgs a) {
  __shared__ double s_cam[FU_WARPS][16];
  __shared__ uint32_t s_stage[FU_WARPS][FU_STAGE];
  __shared__ float4 s_rec[FU_WARPS][96];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t cam = (uint64_t)blockIdx.x * FU_WARPS + warp;
  if (cam >= a.C) return;
  double *c = s_cam[warp];
  uint32_t *stage = s_stage[warp];
  if (lane < 15) c[lane] = __ldg(&a.cams[15 * cam + lane]);
  __syncwarp();
  const V3 cen{a.cen_x[cam], a.cen_y[cam], a.cen_z[cam]};
  const double cc[3] = {cen.x, cen.y, cen.z};
  int lo[3], hi[3];
  if (!camera_cell_range(a.g, cc, a.max_dist, lo, hi)) {
    if (lane == 0) a.vis_count[cam] = 0;
    return;
  }
  const float ox = d2f(cen.x), oy = d2f(cen.y), oz = d2f(cen.z);
  uint32_t n_list = 0;
  const uint32_t *mylist = nullptr;
  if (OCC == FU_OCC_MESH) {
    n_list = a.tri_count[cam];
    mylist = a.tri_list + cam * a.tri_cap;
  }
  uint32_t *out = a.scratch_idx + a.ev_count[cam];
  const uint32_t out_cap = a.ev_count[cam + 1] - a.ev_count[cam];
  uint32_t nvis = 0, found = 0;
  unsigned n_vis_nodes = 0, n_tri = 0;
  int qn = 0;  // warp-uniform number of staged survivors

  auto resolve = [&](int count) {
    const bool have = lane < count;
    Ray ray;
    ray.ox = ox; ray.oy = oy; ray.oz = oz;
    ray.dx = ray.dy = ray.dz = 1.0f;
    ray.tfar = -1.0f;
    uint32_t pt = 0;
    V3 p{0.0, 0.0, 0.0};
    if (have) {
      const uint32_t i = stage[lane];
      p = V3{a.gx[i], a.gy[i], a.gz[i]};
      pt = a.gidx[i];
      if (OCC == FU_OCC_MESH) ray = make_ray(cen, p, a.endpoint_guard_rel != 0);
    }
    bool occ = false;
    if (OCC == FU_OCC_MESH) {
      if (n_list == TRILIST_OVERFLOW)
        occ = warp_any_hit<COUNT>(a.nodes, a.tris, a.n_nodes, ray, have, a.scene_absmax, a.counters);
      else
        occ = packet_any_hit<COUNT>(a, mylist, n_list, ray, have, ox, oy, oz, s_rec[warp], lane, n_vis_nodes, n_tri);
    } else if (OCC == FU_OCC_ANALYTIC) {
      occ = have && hits_building(cen, p, a.block_length, a.block_inset);
    }
    const bool vis = have && !occ;
    const unsigned m = __ballot_sync(0xffffffffu, vis);
    if (vis) {
      const uint32_t pos = nvis + __popc(m & ((1u << lane) - 1u));
      if (pos < out_cap)
        out[pos] = pt;
      else
        a.counters[5] = 1ull;  // cannot happen (visible <= planned row points); reported as an error
    }
    nvis += __popc(m);
  };

  const double r2x = c[2], r2y = c[5], r2z = c[8], tz = c[11];
  for (int z = lo[2]; z <= hi[2]; ++z)
    for (int y = lo[1]; y <= hi[1]; ++y) {
      int x0 = lo[0], x1 = hi[0];
      if (!row_span(a.g, r2x, r2y, r2z, tz, cc[0], a.max_dist, y, z, x0, x1)) continue;
      const uint32_t row = ((uint32_t)z * a.g.n[1] + y) * a.g.n[0];
      const uint32_t start = a.cell_start[row + x0], end = a.cell_start[row + x1 + 1];
      for (uint32_t base = start; base < end; base += 32) {
        const uint32_t i = base + lane;
        bool pass = false;
        if (i < end) {
          double u, v;
          pass = cull_project_thr(c, cen, V3{a.gx[i], a.gy[i], a.gz[i]}, a.t_star, u, v);
        }
        const unsigned m = __ballot_sync(0xffffffffu, pass);
        if (m == 0u) continue;
        if (pass) stage[qn + __popc(m & ((1u << lane) - 1u))] = i;
        qn += __popc(m);
        found += __popc(m);
        __syncwarp();
        if (qn >= 32) {
          resolve(32);
          const int rem = qn - 32;
          uint32_t t = 0;
          if (lane < rem) t = stage[32 + lane];
          __syncwarp();
          if (lane < rem) stage[lane] = t;
          __syncwarp();
          qn = rem;
        }
      }
    }
  if (qn > 0) resolve(qn);
  if (lane == 0) {
    a.vis_count[cam] = nvis;
    if (found) atomicAdd(&a.counters[4], (unsigned long long)found);
    if (COUNT) {
      atomicAdd(&a.counters[2], (unsigned long long)n_vis_nodes);
      atomicAdd(&a.counters[3], (unsigned long long)n_tri);
    }
  }
}

// ---- per-camera sort + final write -------------------------------------------------------------------
// Bitonic network over E*32 keys in registers, element index i = r*32 + lane (so that loads and the
// final stores are coalesced): partner distance j < 32 is one __shfl_xor per register, j >= 32 a
// register-to-register compare.  The (k, j) loops are NOT unrolled — only the E registers of a stage
// are — which keeps the code a few hundred instructions (the fully unrolled network thrashed the
// instruction cache: 8 of 12 stall cycles were no_instruction, profiles/r01c).
template <int E, int RJ>
__device__ __forceinline__ void bitonic_reg_stage(uint32_t (&a)[E], unsigned kr) {
  // partner r ^ RJ; ascending when (r & kr) == 0, kr = k / 32 (kr == E: always ascending)
#pragma unroll
  for (int r = 0; r < E; ++r) {
    const int p = r ^ RJ;
    if (p > r) {
      const bool up = ((unsigned)r & kr) == 0u;
      const uint32_t lo = min(a[r], a[p]), hi = max(a[r], a[p]);
      a[r] = up ? lo : hi;
      a[p] = up ? hi : lo;
    }
  }
}

template <int E>
__device__ __forceinline__ void warp_bitonic_sort_rl(uint32_t (&a)[E], int lane) {
#pragma unroll 1
  for (unsigned k = 2; k <= (unsigned)E * 32u; k <<= 1) {
#pragma unroll 1
    for (unsigned j = k >> 1; j > 0; j >>= 1) {
      if (j < 32u) {
        // i & k: a lane bit when k < 32, a register bit otherwise
        const bool lane_up = ((unsigned)lane & k) == 0u;
        const bool lower = ((unsigned)lane & j) == 0u;
        const unsigned kr = k >> 5;
#pragma unroll
        for (int r = 0; r < E; ++r) {
          const bool up = k < 32u ? lane_up : (((unsigned)r & kr) == 0u);
          const uint32_t other = __shfl_xor_sync(0xffffffffu, a[r], (int)j);
          a[r] = (lower == up) ? min(a[r], other) : max(a[r], other);
        }
      } else {
        const unsigned rj = j >> 5, kr = k >> 5;
        if constexpr (E > 1) { if (rj == 1u) bitonic_reg_stage<E, 1>(a, kr); }
        if constexpr (E > 2) { if (rj == 2u) bitonic_reg_stage<E, 2>(a, kr); }
        if constexpr (E > 4) { if (rj == 4u) bitonic_reg_stage<E, 4>(a, kr); }
        if constexpr (E > 8) { if (rj == 8u) bitonic_reg_stage<E, 8>(a, kr); }
        if constexpr (E > 16) { if (rj == 16u) bitonic_reg_stage<E, 16>(a, kr); }
      }
    }
  }
}

struct SortWriteArgs {
  const uint32_t *ev_off;     // [C+1] scratch slice starts
  const uint32_t *seg_off;    // [C+1] exclusive scan of vis_count = CSR offsets
  uint64_t C;
  const uint32_t *scratch_idx;
  const double *cams;
  const double *p_aos;  // xyz records, original point order
  uint64_t *out_offsets;
  uint64_t *out_idx;
  double2 *out_uv;
};

constexpr uint32_t SW_WARP_MAX = 1024;   // register sort, one warp per camera
constexpr uint32_t SW_BLOCK_MAX = 4096;  // shared-memory sort, one block per camera

constexpr int SW_WARPS = 4;

template <int E>
__device__ __forceinline__ void sort_warp_to_smem(const uint32_t *__restrict__ src, uint32_t n, int lane,
                                                  uint32_t *sorted) {
  uint32_t a[E];
#pragma unroll
  for (int r = 0; r < E; ++r) {
    const uint32_t t = r * 32 + lane;
    a[r] = t < n ? src[t] : 0xffffffffu;
  }
  warp_bitonic_sort_rl<E>(a, lane);
#pragma unroll
  for (int r = 0; r < E; ++r) sorted[r * 32 + lane] = a[r];
}

__global__ void __launch_bounds__(SW_WARPS * 32) k_sort_write(SortWriteArgs s) {
  __shared__ uint32_t s_sorted[SW_WARPS][SW_WARP_MAX];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t cam = (uint64_t)blockIdx.x * SW_WARPS + warp;
  if (cam >= s.C) return;
  const uint32_t base = s.seg_off[cam], n = s.seg_off[cam + 1] - base;
  if (lane == 0) {
    s.out_offsets[cam] = base;
    if (cam == s.C - 1) s.out_offsets[s.C] = s.seg_off[s.C];
  }
  if (n == 0 || n > SW_WARP_MAX) return;
  const uint32_t *src = s.scratch_idx + s.ev_off[cam];
  uint32_t *sorted = s_sorted[warp];
  if (n <= 128)
    sort_warp_to_smem<4>(src, n, lane, sorted);
  else if (n <= 256)
    sort_warp_to_smem<8>(src, n, lane, sorted);
  else if (n <= 512)
    sort_warp_to_smem<16>(src, n, lane, sorted);
  else
    sort_warp_to_smem<32>(src, n, lane, sorted);
  __syncwarp();
  double c[15];
#pragma unroll
  for (int k = 0; k < 15; ++k) c[k] = __ldg(&s.cams[15 * cam + k]);
#pragma unroll 2
  for (uint32_t i = lane; i < n; i += 32) {
    const uint32_t pt = sorted[i];
    const double *p = s.p_aos + 3 * (uint64_t)pt;
    s.out_idx[base + i] = pt;
    s.out_uv[base + i] = observe(c, p[0], p[1], p[2]);
  }
}

// cameras with SW_WARP_MAX < n <= SW_BLOCK_MAX: bitonic network in shared memory, one block each
__global__ void __launch_bounds__(256) k_sort_write_block(SortWriteArgs s) {
  __shared__ uint32_t s_sort[SW_BLOCK_MAX];
  const uint64_t cam = blockIdx.x;
  const uint32_t base = s.seg_off[cam], n = s.seg_off[cam + 1] - base;
  if (n <= SW_WARP_MAX || n > SW_BLOCK_MAX) return;
  const uint32_t *src = s.scratch_idx + s.ev_off[cam];
  uint32_t n2 = 2;
  while (n2 < n) n2 <<= 1;
  for (uint32_t t = threadIdx.x; t < n2; t += blockDim.x) s_sort[t] = t < n ? src[t] : 0xffffffffu;
  __syncthreads();
  for (uint32_t k = 2; k <= n2; k <<= 1) {
    for (uint32_t j = k >> 1; j > 0; j >>= 1) {
      for (uint32_t t = threadIdx.x; t < n2; t += blockDim.x) {
        const uint32_t p = t ^ j;
        if (p > t) {
          const uint32_t x = s_sort[t], y = s_sort[p];
          const bool up = (t & k) == 0;
          if ((x > y) == up) {
            s_sort[t] = y;
            s_sort[p] = x;
          }
        }
      }
      __syncthreads();
    }
  }
  double c[15];
#pragma unroll
  for (int k = 0; k < 15; ++k) c[k] = __ldg(&s.cams[15 * cam + k]);
  for (uint32_t t = threadIdx.x; t < n; t += blockDim.x) {
    const uint32_t pt = s_sort[t];
    const double *p = s.p_aos + 3 * (uint64_t)pt;
    s.out_idx[base + t] = pt;
    s.out_uv[base + t] = observe(c, p[0], p[1], p[2]);
  }
}

// fallback for cameras that see more than SW_BLOCK_MAX points: 64-bit (camera, point) keys of ALL
// visible entries, radix-sorted globally, then k_write_sorted
__global__ void __launch_bounds__(256) k_expand_keys(const uint32_t *__restrict__ ev_off,
                                                     const uint32_t *__restrict__ seg_off, uint64_t C,
                                                     const uint32_t *__restrict__ scratch_idx, int pbits,
                                                     uint64_t *__restrict__ keys) {
  const int lane = threadIdx.x & 31;
  const uint64_t cam = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (cam >= C) return;
  const uint32_t base = seg_off[cam], n = seg_off[cam + 1] - base;
  const uint32_t *src = scratch_idx + ev_off[cam];
  for (uint32_t t = lane; t < n; t += 32) keys[base + t] = (cam << pbits) | (uint64_t)src[t];
}

}  // namespace c2b
