// c2b_noise.cuh — kernel family (4): the `noise` pass (src/noise.rs:35-177) as Philox-keyed
// elementwise kernels, plus the two global statistics it needs (BAProblem::mean/std,
// src/baproblem.rs:282-304, and the element nearest the world origin, src/noise.rs:75-87) as
// deterministic two-stage tree reductions.
//
// Random stream: Philox4x32-10, key = seed, counter = (index lo, index hi, stream, slot); each
// counter yields two N(0,1) draws by Box-Muller in f64 (u1 in (0,1], u2 in [0,1)).  Streams and
// slots follow the reference's draw order (angle before translation, axis before magnitude).
#pragma once
#include "c2b_common.cuh"
#include "c2b_math.cuh"

namespace c2b {

enum { ST_DRIFT_CAM = 1, ST_DRIFT_PT = 2, ST_NOISE_CAM = 3, ST_NOISE_PT = 4, ST_NOISE_OBS = 5 };

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t *out) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0;
    c1 = lo1;
    c2 = n2;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0;
  out[1] = c1;
  out[2] = c2;
  out[3] = c3;
}

__device__ __forceinline__ void normal_pair(uint64_t seed, uint32_t stream, uint64_t index,
                                            uint32_t slot, double &z0, double &z1) {
  uint32_t o[4];
  philox4x32_10((uint32_t)index, (uint32_t)(index >> 32), stream, slot, (uint32_t)seed,
                (uint32_t)(seed >> 32), o);
  uint64_t x = (uint64_t)o[0] | ((uint64_t)o[1] << 32), y = (uint64_t)o[2] | ((uint64_t)o[3] << 32);
  double u1 = dmul((double)((x >> 11) + 1), 1.1102230246251565e-16);
  double u2 = dmul((double)(y >> 11), 1.1102230246251565e-16);
  double rr = dsqrt(dmul(-2.0, log(u1)));
  double th = dmul(6.283185307179586, u2);
  double s, c;
  sincos(th, &s, &c);
  z0 = dmul(rr, c);
  z1 = dmul(rr, s);
}

// element i of the chained sequence "camera centres, then points" (src/baproblem.rs:284-287)
__device__ __forceinline__ V3 chain_element(const double *cx, const double *cy, const double *cz,
                                            uint64_t C, const double *pts, uint64_t i) {
  if (i < C) return V3{cx[i], cy[i], cz[i]};
  const double *p = pts + 3 * (i - C);
  return V3{p[0], p[1], p[2]};
}

constexpr int ST_BLOCKS = 592;  // 4 x 148 SMs
constexpr int ST_THREADS = 256;

// MODE 0: sum of e/num ; MODE 1: sum of (e-mean)^2 ; partial[3*block + k]
template <int MODE>
__global__ void __launch_bounds__(ST_THREADS)
    k_stats_partial(const double *__restrict__ cx, const double *__restrict__ cy,
                    const double *__restrict__ cz, uint64_t C, const double *__restrict__ pts,
                    uint64_t P, double num, V3 mean, double *__restrict__ partial) {
  double s[3] = {0, 0, 0};
  uint64_t n = C + P;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x) {
    V3 e = chain_element(cx, cy, cz, C, pts, i);
    if (MODE == 0) {
      s[0] += e.x / num;
      s[1] += e.y / num;
      s[2] += e.z / num;
    } else {
      double a = e.x - mean.x, b = e.y - mean.y, c = e.z - mean.z;
      s[0] += a * a;
      s[1] += b * b;
      s[2] += c * c;
    }
  }
  __shared__ double sh[3][ST_THREADS / 32];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
    if ((threadIdx.x & 31) == 0) sh[k][threadIdx.x >> 5] = s[k];
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double r = 0;
    for (int w = 0; w < ST_THREADS / 32; ++w) r += sh[threadIdx.x][w];
    partial[3 * (uint64_t)blockIdx.x + threadIdx.x] = r;
  }
}

__global__ void k_stats_final(const double *__restrict__ partial, int nb, double *__restrict__ out3) {
  if (threadIdx.x < 3) {
    double r = 0;
    for (int b = 0; b < nb; ++b) r += partial[3 * b + threadIdx.x];
    out3[threadIdx.x] = r;
  }
}

// nearest-to-origin element; the reference's fold keeps the LATER element on ties
// (src/noise.rs:80-86), so the reduction orders by (distance asc, index desc).
__device__ __forceinline__ bool nearer(double da, uint64_t ia, double db, uint64_t ib) {
  return da < db || (da == db && ia > ib);
}
__global__ void __launch_bounds__(ST_THREADS)
    k_nearest_partial(const double *__restrict__ cx, const double *__restrict__ cy,
                      const double *__restrict__ cz, uint64_t C, const double *__restrict__ pts,
                      uint64_t P, double *__restrict__ pd, unsigned long long *__restrict__ pi) {
  double bd = INFINITY;
  uint64_t bi = 0;
  bool have = false;
  uint64_t n = C + P;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x) {
    double d = mag(chain_element(cx, cy, cz, C, pts, i));
    if (!have || nearer(d, i, bd, bi) || !(bd == bd)) {
      bd = d;
      bi = i;
      have = true;
    }
  }
  if (!have) {
    bd = INFINITY;
    bi = 0;
  }
  __shared__ double sd[ST_THREADS];
  __shared__ unsigned long long si[ST_THREADS];
  __shared__ int sv[ST_THREADS];
  sd[threadIdx.x] = bd;
  si[threadIdx.x] = bi;
  sv[threadIdx.x] = have ? 1 : 0;
  __syncthreads();
  for (int o = ST_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      int a = threadIdx.x, b = threadIdx.x + o;
      if (sv[b] && (!sv[a] || nearer(sd[b], si[b], sd[a], si[a]))) {
        sd[a] = sd[b];
        si[a] = si[b];
        sv[a] = 1;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    pd[blockIdx.x] = sv[0] ? sd[0] : INFINITY;
    pi[blockIdx.x] = sv[0] ? si[0] : ~0ull;
  }
}
__global__ void k_nearest_final(const double *__restrict__ pd, const unsigned long long *__restrict__ pi,
                                int nb, const double *__restrict__ cx, const double *__restrict__ cy,
                                const double *__restrict__ cz, uint64_t C,
                                const double *__restrict__ pts, double *__restrict__ origin3) {
  if (threadIdx.x == 0) {
    double bd = INFINITY;
    unsigned long long bi = ~0ull;
    for (int b = 0; b < nb; ++b) {
      if (pi[b] == ~0ull) continue;
      if (bi == ~0ull || nearer(pd[b], pi[b], bd, bi)) {
        bd = pd[b];
        bi = pi[b];
      }
    }
    V3 e = bi == ~0ull ? V3{0, 0, 0} : chain_element(cx, cy, cz, C, pts, bi);
    origin3[0] = e.x;
    origin3[1] = e.y;
    origin3[2] = e.z;
  }
}

// ---- add_drift, src/noise.rs:68-116 ------------------------------------------------------------------
__global__ void k_drift_cams(double *__restrict__ cams, uint64_t C, const double *__restrict__ origin,
                             V3 dir, double strength, double angle_strength, double std,
                             uint64_t seed) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C) return;
  double cam[15];
#pragma unroll
  for (int k = 0; k < 15; ++k) cam[k] = cams[15 * i + k];
  V3 c = camera_center(cam);
  V3 d{dsub(c.x, origin[0]), dsub(c.y, origin[1]), dsub(c.z, origin[2])};
  double distance = mag(d);
  double z0 = 0.0, z1 = 0.0;
  if (std != 0.0) normal_pair(seed, ST_DRIFT_CAM, i, 0, z0, z1);  // Normal(1, 0) is 1 whatever the draw
  double v1 = dadd(1.0, dmul(std, z0)), v2 = dadd(1.0, dmul(std, z1));
  double angle = dmul(dmul(angle_strength, v1), pow(distance, 1.2));
  V3 dl{dmul(dmul(dmul(dmul(dir.x, strength), v2), distance), distance),
        dmul(dmul(dmul(dmul(dir.y, strength), v2), distance), distance),
        dmul(dmul(dmul(dmul(dir.z, strength), v2), distance), distance)};
  double R[9], out[15];
  from_angle_x(angle, R);
  camera_transform(cam, R, dl, out);
#pragma unroll
  for (int k = 0; k < 15; ++k) cams[15 * i + k] = out[k];
}

__global__ void k_drift_pts(double *__restrict__ pts, uint64_t P, const double *__restrict__ origin,
                            V3 dir, double strength, double std, uint64_t seed) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  V3 p{pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]};
  V3 d{dsub(p.x, origin[0]), dsub(p.y, origin[1]), dsub(p.z, origin[2])};
  double distance = mag(d);
  double z0 = 0.0, z1 = 0.0;
  if (std != 0.0) normal_pair(seed, ST_DRIFT_PT, i, 0, z0, z1);  // Normal(1, 0) is 1 whatever the draw
  double v = dadd(1.0, dmul(std, z0));
  pts[3 * i] = dadd(p.x, dmul(dmul(dmul(dmul(dir.x, strength), v), distance), distance));
  pts[3 * i + 1] = dadd(p.y, dmul(dmul(dmul(dmul(dir.y, strength), v), distance), distance));
  pts[3 * i + 2] = dadd(p.z, dmul(dmul(dmul(dmul(dir.z, strength), v), distance), distance));
}

// ---- add_noise, src/noise.rs:119-177 --------------------------------------------------------------------
__global__ void k_noise_cams(double *__restrict__ cams, uint64_t C, double bal_std,
                             double translation_std, double rotation_std, uint64_t seed) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C) return;
  double cam[15];
#pragma unroll
  for (int k = 0; k < 15; ++k) cam[k] = cams[15 * i + k];
  double a0, a1, a2, ang, b0, b1, b2, mg;
  normal_pair(seed, ST_NOISE_CAM, i, 0, a0, a1);
  normal_pair(seed, ST_NOISE_CAM, i, 1, a2, ang);
  normal_pair(seed, ST_NOISE_CAM, i, 2, b0, b1);
  normal_pair(seed, ST_NOISE_CAM, i, 3, b2, mg);
  V3 ax = normalize(V3{a0, a1, a2});
  double angle = dadd(0.0, dmul(rotation_std, ang));
  V3 tr = normalize(V3{b0, b1, b2});
  double m = dadd(0.0, dmul(translation_std, mg));
  double R[9], out[15];
  from_axis_angle(ax, angle, R);
  V3 dl{dmul(dmul(tr.x, bal_std), m), dmul(dmul(tr.y, bal_std), m), dmul(dmul(tr.z, bal_std), m)};
  camera_transform(cam, R, dl, out);
#pragma unroll
  for (int k = 0; k < 15; ++k) cams[15 * i + k] = out[k];
}

__global__ void k_noise_pts(double *__restrict__ pts, uint64_t P, double point_std, uint64_t seed) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  double a0, a1, a2, mg;
  normal_pair(seed, ST_NOISE_PT, i, 0, a0, a1);
  normal_pair(seed, ST_NOISE_PT, i, 1, a2, mg);
  V3 ax = normalize(V3{a0, a1, a2});
  double m = dadd(0.0, dmul(point_std, mg));
  pts[3 * i] = dadd(pts[3 * i], dmul(ax.x, m));
  pts[3 * i + 1] = dadd(pts[3 * i + 1], dmul(ax.y, m));
  pts[3 * i + 2] = dadd(pts[3 * i + 2], dmul(ax.z, m));
}

__global__ void k_noise_obs(double2 *__restrict__ uv, uint64_t O, double observations_std,
                            uint64_t seed) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= O) return;
  double nx, ny, r0, unused;
  normal_pair(seed, ST_NOISE_OBS, i, 0, nx, ny);
  normal_pair(seed, ST_NOISE_OBS, i, 1, r0, unused);
  double m = dsqrt(dadd(dmul(nx, nx), dmul(ny, ny)));
  double r = dadd(0.0, dmul(observations_std, r0));
  double2 q = uv[i];
  q.x = dadd(q.x, dmul(ddiv(nx, m), r));
  q.y = dadd(q.y, dmul(ddiv(ny, m), r));
  uv[i] = q;
}

// ---- add_sin_noise, src/noise.rs:388-416 --------------------------------------------------------------
// extent of the chained sequence (BAProblem::extent, src/baproblem.rs:307-330): partial[6*block + k]
__global__ void __launch_bounds__(ST_THREADS)
    k_extent_partial(const double *__restrict__ cx, const double *__restrict__ cy,
                     const double *__restrict__ cz, uint64_t C, const double *__restrict__ pts,
                     uint64_t P, double *__restrict__ partial) {
  double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  uint64_t n = C + P;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x) {
    V3 e = chain_element(cx, cy, cz, C, pts, i);
    // f64::min / f64::max semantics (a NaN operand is ignored), like fmin / fmax
    lo[0] = fmin(lo[0], e.x); lo[1] = fmin(lo[1], e.y); lo[2] = fmin(lo[2], e.z);
    hi[0] = fmax(hi[0], e.x); hi[1] = fmax(hi[1], e.y); hi[2] = fmax(hi[2], e.z);
  }
  __shared__ double sh[6][ST_THREADS / 32];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo[k] = fmin(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
      hi[k] = fmax(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
    }
    if ((threadIdx.x & 31) == 0) {
      sh[k][threadIdx.x >> 5] = lo[k];
      sh[3 + k][threadIdx.x >> 5] = hi[k];
    }
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    double r = sh[threadIdx.x][0];
    for (int w = 1; w < ST_THREADS / 32; ++w)
      r = threadIdx.x < 3 ? fmin(r, sh[threadIdx.x][w]) : fmax(r, sh[threadIdx.x][w]);
    partial[6 * (uint64_t)blockIdx.x + threadIdx.x] = r;
  }
}
// out6 = min xyz, max xyz
__global__ void k_extent_final(const double *__restrict__ partial, int nb, double *__restrict__ out6) {
  if (threadIdx.x < 6) {
    double r = partial[threadIdx.x];
    for (int b = 1; b < nb; ++b)
      r = threadIdx.x < 3 ? fmin(r, partial[6 * b + threadIdx.x]) : fmax(r, partial[6 * b + threadIdx.x]);
    out6[threadIdx.x] = r;
  }
}

// noise(x) = sin(dot(x / dimension, dir) * frequency * pi) * strength * normalize(noise_dir)
__device__ __forceinline__ V3 sin_noise(V3 x, V3 dim, V3 dir, V3 nd, double strength, double frequency) {
  const V3 q{ddiv(x.x, dim.x), ddiv(x.y, dim.y), ddiv(x.z, dim.z)};
  const double dot = dadd(dadd(dmul(q.x, dir.x), dmul(q.y, dir.y)), dmul(q.z, dir.z));
  const double amp = dmul(sin(dmul(dmul(dot, frequency), 3.14159265358979323846)), strength);
  return V3{dmul(nd.x, amp), dmul(nd.y, amp), dmul(nd.z, amp)};
}

__global__ void k_sin_cams(double *__restrict__ cams, uint64_t C, V3 dim, V3 dir, V3 nd, double strength,
                           double frequency) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C) return;
  double cam[15], out[15];
#pragma unroll
  for (int k = 0; k < 15; ++k) cam[k] = cams[15 * i + k];
  const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};  // Basis3::one()
  camera_transform(cam, I, sin_noise(camera_center(cam), dim, dir, nd, strength, frequency), out);
#pragma unroll
  for (int k = 0; k < 15; ++k) cams[15 * i + k] = out[k];
}

__global__ void k_sin_pts(double *__restrict__ pts, uint64_t P, V3 dim, V3 dir, V3 nd, double strength,
                          double frequency) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const V3 p{pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]};
  const V3 d = sin_noise(p, dim, dir, nd, strength, frequency);
  pts[3 * i] = dadd(p.x, d.x);
  pts[3 * i + 1] = dadd(p.y, d.y);
  pts[3 * i + 2] = dadd(p.z, d.z);
}

// ---- total_reprojection_error, src/baproblem.rs:265-279 (one thread per observation, two-stage sum)
__global__ void __launch_bounds__(ST_THREADS)
    k_reproj_partial(const double *__restrict__ cams, const double *__restrict__ px,
                     const double *__restrict__ py, const double *__restrict__ pz,
                     const uint64_t *__restrict__ offsets, uint64_t C,
                     const uint32_t *__restrict__ idx, const double2 *__restrict__ uv, uint64_t O,
                     double norm, double *__restrict__ partial) {
  double s = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < O;
       i += (uint64_t)gridDim.x * blockDim.x) {
    // camera of observation i: largest c with offsets[c] <= i
    uint64_t lo = 0, hi = C;
    while (hi - lo > 1) {
      uint64_t mid = (lo + hi) >> 1;
      if (offsets[mid] <= i) lo = mid; else hi = mid;
    }
    double cam[15];
#pragma unroll
    for (int k = 0; k < 15; ++k) cam[k] = __ldg(&cams[15 * lo + k]);
    uint64_t pt = idx[i];
    V3 pc = project_world(cam, V3{px[pt], py[pt], pz[pt]});
    double u, v;
    project(cam[12], cam[13], cam[14], pc, u, v);
    double du = fabs(u - uv[i].x), dv = fabs(v - uv[i].y);
    if (norm == 1.0) s += du + dv;
    else if (norm == 2.0) s += du * du + dv * dv;
    else s += pow(du, norm) + pow(dv, norm);
  }
  __shared__ double sh[ST_THREADS / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double r = 0;
    for (int w = 0; w < ST_THREADS / 32; ++w) r += sh[w];
    partial[blockIdx.x] = r;
  }
}

}  // namespace c2b
