// c2b_math.cuh — SnavelyCamera arithmetic and the ray/triangle predicate, written once for
// host and device with every rounding step explicit.
//
// The reference is Rust (no FMA contraction, IEEE round-to-nearest f64/f32).  On the device
// each operation below is a __d*_rn / __f*_rn intrinsic, which nvcc never fuses; on the host
// the translation unit is compiled with -ffp-contract=off.  The operation ORDER follows cgmath
// 0.17 as used by src/baproblem.rs:141-176 (Matrix3*Vector3 = col0*x + col1*y + col2*z, left to
// right; Basis3::invert = general cofactor inverse; normalize = v * (1/|v|)).
#pragma once
#include <cstdint>
#include <cmath>

#if defined(__CUDACC__)
#define C2B_HD __host__ __device__ __forceinline__
#else
#define C2B_HD inline
#endif

namespace c2b {

#if defined(__CUDA_ARCH__)
C2B_HD double dmul(double a, double b) { return __dmul_rn(a, b); }
C2B_HD double dadd(double a, double b) { return __dadd_rn(a, b); }
C2B_HD double dsub(double a, double b) { return __dsub_rn(a, b); }
C2B_HD double ddiv(double a, double b) { return __ddiv_rn(a, b); }
C2B_HD double dsqrt(double a) { return __dsqrt_rn(a); }
C2B_HD float fmul(float a, float b) { return __fmul_rn(a, b); }
C2B_HD float fadd(float a, float b) { return __fadd_rn(a, b); }
C2B_HD float fsub(float a, float b) { return __fsub_rn(a, b); }
C2B_HD float fdiv(float a, float b) { return __fdiv_rn(a, b); }
C2B_HD float d2f(double a) { return __double2float_rn(a); }
#else
C2B_HD double dmul(double a, double b) { return a * b; }
C2B_HD double dadd(double a, double b) { return a + b; }
C2B_HD double dsub(double a, double b) { return a - b; }
C2B_HD double ddiv(double a, double b) { return a / b; }
C2B_HD double dsqrt(double a) { return sqrt(a); }
C2B_HD float fmul(float a, float b) { return a * b; }
C2B_HD float fadd(float a, float b) { return a + b; }
C2B_HD float fsub(float a, float b) { return a - b; }
C2B_HD float fdiv(float a, float b) { return a / b; }
C2B_HD float d2f(double a) { return (float)a; }
#endif

struct V3 {
  double x, y, z;
};

// R is column-major: R[0..2] = column x.  (cgmath Matrix3 * Vector3)
C2B_HD V3 mat_vec(const double *R, V3 v) {
  V3 o;
  o.x = dadd(dadd(dmul(R[0], v.x), dmul(R[3], v.y)), dmul(R[6], v.z));
  o.y = dadd(dadd(dmul(R[1], v.x), dmul(R[4], v.y)), dmul(R[7], v.z));
  o.z = dadd(dadd(dmul(R[2], v.x), dmul(R[5], v.y)), dmul(R[8], v.z));
  return o;
}

C2B_HD V3 cross(V3 a, V3 b) {
  V3 o;
  o.x = dsub(dmul(a.y, b.z), dmul(a.z, b.y));
  o.y = dsub(dmul(a.z, b.x), dmul(a.x, b.z));
  o.z = dsub(dmul(a.x, b.y), dmul(a.y, b.x));
  return o;
}

C2B_HD double mag2(V3 v) { return dadd(dadd(dmul(v.x, v.x), dmul(v.y, v.y)), dmul(v.z, v.z)); }
C2B_HD double mag(V3 v) { return dsqrt(mag2(v)); }

// general 3x3 inverse (cgmath Matrix3::invert), column-major in and out
C2B_HD void mat_invert(const double *M, double *O) {
  V3 c0{M[0], M[1], M[2]}, c1{M[3], M[4], M[5]}, c2{M[6], M[7], M[8]};
  double det = dadd(dsub(dmul(c0.x, dsub(dmul(c1.y, c2.z), dmul(c2.y, c1.z))),
                         dmul(c1.x, dsub(dmul(c0.y, c2.z), dmul(c2.y, c0.z)))),
                    dmul(c2.x, dsub(dmul(c0.y, c1.z), dmul(c1.y, c0.z))));
  V3 r0 = cross(c1, c2), r1 = cross(c2, c0), r2 = cross(c0, c1);
  O[0] = ddiv(r0.x, det);
  O[1] = ddiv(r1.x, det);
  O[2] = ddiv(r2.x, det);
  O[3] = ddiv(r0.y, det);
  O[4] = ddiv(r1.y, det);
  O[5] = ddiv(r2.y, det);
  O[6] = ddiv(r0.z, det);
  O[7] = ddiv(r1.z, det);
  O[8] = ddiv(r2.z, det);
}

// SnavelyCamera::center, src/baproblem.rs:161-163
C2B_HD V3 camera_center(const double *cam) {
  double inv[9];
  mat_invert(cam, inv);
  V3 r = mat_vec(inv, V3{cam[9], cam[10], cam[11]});
  return V3{-r.x, -r.y, -r.z};
}

// SnavelyCamera::project_world, src/baproblem.rs:141-143
C2B_HD V3 project_world(const double *cam, V3 p) {
  V3 r = mat_vec(cam, p);
  return V3{dadd(r.x, cam[9]), dadd(r.y, cam[10]), dadd(r.z, cam[11])};
}

// a / b of the perspective division, bit for bit IEEE 754 like ddiv.  The device's division routine sends a
// ZERO NUMERATOR through its slow path (a call of ~55 dependent instructions), and the reference's synthetic
// cities produce exactly that for every observation: cameras and points share one height and the camera axes
// are lattice-aligned, so pc.y == 0 exactly.  At cfg4 that call was 24 % of k_sort_write's warp instructions
// (profiles/r02j_sass_dynamic_sort_write.txt).  (+-0) / b = +-0 with sign(a) ^ sign(b) for every b that is neither zero
// nor NaN (infinite b included); everything else takes the general division.
C2B_HD double ddiv_persp(double a, double b) {
#if defined(__CUDA_ARCH__)
  if (a == 0.0 && b == b && b != 0.0)
    return __hiloint2double((__double2hiint(a) ^ __double2hiint(b)) & (int)0x80000000, 0);
#endif
  return ddiv(a, b);
}

// SnavelyCamera::project, src/baproblem.rs:145-151; |p_|^4 as m2*m2
C2B_HD void project(double f, double k1, double k2, V3 pc, double &u, double &v) {
  double px = ddiv_persp(-pc.x, pc.z), py = ddiv_persp(-pc.y, pc.z);
  double m2 = dadd(dmul(px, px), dmul(py, py));
  double r = dadd(dadd(1.0, dmul(k1, m2)), dmul(k2, dmul(m2, m2)));
  double fr = dmul(f, r);
  u = dmul(fr, px);
  v = dmul(fr, py);
}

// the cull + projection predicate of src/generate.rs:448-454 / src/synthetic.rs:285-289
C2B_HD bool cull_project(const double *cam, V3 center, V3 p, double max_dist, double &u,
                         double &v) {
  V3 pc = project_world(cam, p);
  V3 d{dsub(center.x, p.x), dsub(center.y, p.y), dsub(center.z, p.z)};
  double dist = mag(d);
  if (!(dist < max_dist && pc.z <= 0.0)) return false;
  project(cam[12], cam[13], cam[14], pc, u, v);
  return u >= -1.0 && u <= 1.0 && v >= -1.0 && v <= 1.0;
}


// SnavelyCamera::from_position_direction, src/baproblem.rs:153-159
C2B_HD void camera_from_position_direction(V3 pos, const double *R, double *cam) {
  V3 r = mat_vec(R, pos);
  for (int k = 0; k < 9; ++k) cam[k] = R[k];
  cam[9] = dmul(-1.0, r.x);
  cam[10] = dmul(-1.0, r.y);
  cam[11] = dmul(-1.0, r.z);
  cam[12] = 1.0;
  cam[13] = 0.0;
  cam[14] = 0.0;
}

// SnavelyCamera::transform, src/baproblem.rs:165-171: dir' = dir * delta_dir,
// loc' = -(dir_OLD * (center + delta_loc)).  out may alias cam.
C2B_HD void camera_transform(const double *cam, const double *dR, V3 dloc, double *out) {
  V3 c = camera_center(cam);
  V3 q{dadd(c.x, dloc.x), dadd(c.y, dloc.y), dadd(c.z, dloc.z)};
  V3 r = mat_vec(cam, q);
  double Rn[9];
  for (int col = 0; col < 3; ++col) {
    V3 o = mat_vec(cam, V3{dR[3 * col], dR[3 * col + 1], dR[3 * col + 2]});
    Rn[3 * col] = o.x;
    Rn[3 * col + 1] = o.y;
    Rn[3 * col + 2] = o.z;
  }
  double i0 = cam[12], i1 = cam[13], i2 = cam[14];
  for (int k = 0; k < 9; ++k) out[k] = Rn[k];
  out[9] = dmul(-1.0, r.x);
  out[10] = dmul(-1.0, r.y);
  out[11] = dmul(-1.0, r.z);
  out[12] = i0;
  out[13] = i1;
  out[14] = i2;
}

// cgmath Matrix3::from_angle_x (column-major)
C2B_HD void from_angle_x(double rad, double *R) {
  double s = sin(rad), c = cos(rad);
  R[0] = 1.0; R[1] = 0.0; R[2] = 0.0;
  R[3] = 0.0; R[4] = c;   R[5] = s;
  R[6] = 0.0; R[7] = -s;  R[8] = c;
}
// cgmath Matrix3::from_angle_y (column-major)
C2B_HD void from_angle_y(double rad, double *R) {
  double s = sin(rad), c = cos(rad);
  R[0] = c;   R[1] = 0.0; R[2] = -s;
  R[3] = 0.0; R[4] = 1.0; R[5] = 0.0;
  R[6] = s;   R[7] = 0.0; R[8] = c;
}
// cgmath Matrix3::from_axis_angle (column-major)
C2B_HD void from_axis_angle(V3 a, double rad, double *R) {
  double s = sin(rad), c = cos(rad);
  double k = dsub(1.0, c);
  R[0] = dadd(dmul(dmul(k, a.x), a.x), c);
  R[1] = dadd(dmul(dmul(k, a.x), a.y), dmul(s, a.z));
  R[2] = dsub(dmul(dmul(k, a.x), a.z), dmul(s, a.y));
  R[3] = dsub(dmul(dmul(k, a.x), a.y), dmul(s, a.z));
  R[4] = dadd(dmul(dmul(k, a.y), a.y), c);
  R[5] = dadd(dmul(dmul(k, a.y), a.z), dmul(s, a.x));
  R[6] = dadd(dmul(dmul(k, a.x), a.z), dmul(s, a.y));
  R[7] = dsub(dmul(dmul(k, a.y), a.z), dmul(s, a.x));
  R[8] = dadd(dmul(dmul(k, a.z), a.z), c);
}

// cgmath InnerSpace::normalize = v * (1/|v|)
C2B_HD V3 normalize(V3 v) {
  double s = ddiv(1.0, mag(v));
  return V3{dmul(v.x, s), dmul(v.y, s), dmul(v.z, s)};
}

// ---- ray construction, src/generate.rs:456-464 -------------------------------------------------
struct Ray {
  float ox, oy, oz;
  float dx, dy, dz;
  float tfar;
};

C2B_HD Ray make_ray(V3 c, V3 p, bool endpoint_guard_rel) {
  V3 d{dsub(p.x, c.x), dsub(p.y, c.y), dsub(p.z, c.z)};
  double n = mag(d);
  double s = ddiv(1.0, n);
  Ray r;
  r.ox = d2f(c.x);
  r.oy = d2f(c.y);
  r.oz = d2f(c.z);
  r.dx = d2f(dmul(d.x, s));
  r.dy = d2f(dmul(d.y, s));
  r.dz = d2f(dmul(d.z, s));
  r.tfar = fsub(d2f(n), 1e-6f);
  if (endpoint_guard_rel) r.tfar = fmul(r.tfar, fsub(1.0f, 3.814697265625e-06f));
  return r;
}

// ---- watertight ray/triangle test, hit interval 0 < t <= tfar ------------------------------------
// Edge functions in the form of Embree's robust ("Pluecker") triangle intersector: with the
// ORIGIN-RELATIVE vertices A = v0 - o, B = v1 - o, C = v2 - o and the edges e0 = C - A,
// e1 = A - B, e2 = B - C,
//     U = d . (e0 x (C + A)),  V = d . (e1 x (A + B)),  W = d . (e2 x (B + C)),
//     det = U + V + W  (= 2 d.N),   T = 2 A.N  with  N = e0 x e1,   t = T / det.
// e x (sum) equals twice the plain cross product of the two vertices but keeps the error
// proportional to the EDGE length, so small triangles far from the origin stay accurate.
// The three edge normals and T depend on (origin, triangle) only: for the rays of one camera they
// are computed once per triangle (TriRec) and a ray test is 9 FMA and a few compares.
// Watertight: differences and sums are exactly anti-/symmetric in their operands and the cross
// products are evaluated unfused (p1 - p2), so the two triangles sharing an edge get edge normals
// that are exact negatives of each other, hence edge-function values of equal magnitude and
// opposite sign; a zero counts as inside for both.  No back-face culling.  NaN anywhere => no hit.
struct TriRec {
  float ux, uy, uz;  // e0 x (C + A)
  float vx, vy, vz;  // e1 x (A + B)
  float wx, wy, wz;  // e2 x (B + C)
  float T;           // 2 A . (e0 x e1)
};

#if defined(__CUDA_ARCH__)
C2B_HD float ffma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
#else
C2B_HD float ffma(float a, float b, float c) { return fmaf(a, b, c); }
#endif

C2B_HD float dot3f(float ax, float ay, float az, float bx, float by, float bz) {
  return ffma(az, bz, ffma(ay, by, fmul(ax, bx)));
}

C2B_HD TriRec tri_record(float ox, float oy, float oz, float v0x, float v0y, float v0z, float v1x,
                         float v1y, float v1z, float v2x, float v2y, float v2z) {
  const float Ax = fsub(v0x, ox), Ay = fsub(v0y, oy), Az = fsub(v0z, oz);
  const float Bx = fsub(v1x, ox), By = fsub(v1y, oy), Bz = fsub(v1z, oz);
  const float Cx = fsub(v2x, ox), Cy = fsub(v2y, oy), Cz = fsub(v2z, oz);
  const float e0x = fsub(Cx, Ax), e0y = fsub(Cy, Ay), e0z = fsub(Cz, Az);
  const float e1x = fsub(Ax, Bx), e1y = fsub(Ay, By), e1z = fsub(Az, Bz);
  const float e2x = fsub(Bx, Cx), e2y = fsub(By, Cy), e2z = fsub(Bz, Cz);
  const float sux = fadd(Cx, Ax), suy = fadd(Cy, Ay), suz = fadd(Cz, Az);
  const float svx = fadd(Ax, Bx), svy = fadd(Ay, By), svz = fadd(Az, Bz);
  const float swx = fadd(Bx, Cx), swy = fadd(By, Cy), swz = fadd(Bz, Cz);
  TriRec t;
  t.ux = fsub(fmul(e0y, suz), fmul(e0z, suy));
  t.uy = fsub(fmul(e0z, sux), fmul(e0x, suz));
  t.uz = fsub(fmul(e0x, suy), fmul(e0y, sux));
  t.vx = fsub(fmul(e1y, svz), fmul(e1z, svy));
  t.vy = fsub(fmul(e1z, svx), fmul(e1x, svz));
  t.vz = fsub(fmul(e1x, svy), fmul(e1y, svx));
  t.wx = fsub(fmul(e2y, swz), fmul(e2z, swy));
  t.wy = fsub(fmul(e2z, swx), fmul(e2x, swz));
  t.wz = fsub(fmul(e2x, swy), fmul(e2y, swx));
  const float nx = fsub(fmul(e0y, e1z), fmul(e0z, e1y));
  const float ny = fsub(fmul(e0z, e1x), fmul(e0x, e1z));
  const float nz = fsub(fmul(e0x, e1y), fmul(e0y, e1x));
  t.T = fmul(2.0f, dot3f(Ax, Ay, Az, nx, ny, nz));
  return t;
}

// true when the triangle occludes the ray (dx, dy, dz, tfar) whose origin the record was built
// for.  t_out (optional) receives t for the closest-hit entry (c2b_intersect1).
C2B_HD bool ray_tri_record(float dx, float dy, float dz, float tfar, const TriRec &t,
                           float *t_out = nullptr) {
  const float U = dot3f(dx, dy, dz, t.ux, t.uy, t.uz);
  const float V = dot3f(dx, dy, dz, t.vx, t.vy, t.vz);
  const float W = dot3f(dx, dy, dz, t.wx, t.wy, t.wz);
  // one boolean expression without early exits (they cost a convergence barrier per test on the GPU):
  // miss if the edge functions have mixed signs or det == 0; NaN anywhere fails the last compare
  const bool mixed = ((U < 0.0f) | (V < 0.0f) | (W < 0.0f)) & ((U > 0.0f) | (V > 0.0f) | (W > 0.0f));
  const float det = fadd(fadd(U, V), W);
  const float ad = fabsf(det);
  const float Ts = det < 0.0f ? -t.T : t.T;
  const bool hit = !mixed & (det != 0.0f) & (Ts > 0.0f) & (Ts <= fmul(tfar, ad));
  if (t_out && hit) *t_out = fdiv(Ts, ad);
  return hit;
}

C2B_HD bool ray_triangle(const Ray &r, float v0x, float v0y, float v0z, float v1x, float v1y,
                         float v1z, float v2x, float v2y, float v2z, float *t_out = nullptr) {
  const TriRec t = tri_record(r.ox, r.oy, r.oz, v0x, v0y, v0z, v1x, v1y, v1z, v2x, v2y, v2z);
  return ray_tri_record(r.dx, r.dy, r.dz, r.tfar, t, t_out);
}

// ---- Moeller-Trumbore test in the form of Embree 3's DEFAULT (non-robust) triangle intersector ----------
// What the reference's scene actually runs (a default scene, src/bin/city2ba.rs:515-521, src/generate.rs:472).
// With e1 = v0 - v1, e2 = v2 - v0, Ng = e2 x e1, C = v0 - org, R = C x dir, den = Ng . dir:
//     hit  <=>  den != 0  and  U >= 0  and  V >= 0  and  U + V <= |den|,  U = (R . e2) ^ sign(den),
//               V = (R . e1) ^ sign(den),  and  0 < T <= |den| tfar  with  T = (Ng . C) ^ sign(den),
// cross products unfused, dot products as fused multiply-add chains a0 b0 + (a1 b1 + a2 b2).  Restated from
// the published algorithm (Embree is not vendored by the reference: unpinned like everything at that
// boundary); not watertight on shared edges.  Selected with c2b_vis_options::predicate = C2B_PRED_MT.
C2B_HD float dot3m(float ax, float ay, float az, float bx, float by, float bz) {
  return ffma(ax, bx, ffma(ay, by, fmul(az, bz)));
}

C2B_HD bool ray_triangle_mt(const Ray &r, float v0x, float v0y, float v0z, float v1x, float v1y, float v1z,
                            float v2x, float v2y, float v2z) {
  const float e1x = fsub(v0x, v1x), e1y = fsub(v0y, v1y), e1z = fsub(v0z, v1z);
  const float e2x = fsub(v2x, v0x), e2y = fsub(v2y, v0y), e2z = fsub(v2z, v0z);
  const float cx = fsub(v0x, r.ox), cy = fsub(v0y, r.oy), cz = fsub(v0z, r.oz);
  const float ngx = fsub(fmul(e2y, e1z), fmul(e2z, e1y));
  const float ngy = fsub(fmul(e2z, e1x), fmul(e2x, e1z));
  const float ngz = fsub(fmul(e2x, e1y), fmul(e2y, e1x));
  const float rx = fsub(fmul(cy, r.dz), fmul(cz, r.dy));
  const float ry = fsub(fmul(cz, r.dx), fmul(cx, r.dz));
  const float rz = fsub(fmul(cx, r.dy), fmul(cy, r.dx));
  const float den = dot3m(ngx, ngy, ngz, r.dx, r.dy, r.dz);
  const float ad = fabsf(den);
  const float sg = den < 0.0f ? -1.0f : 1.0f;
  const float U = fmul(dot3m(rx, ry, rz, e2x, e2y, e2z), sg), V = fmul(dot3m(rx, ry, rz, e1x, e1y, e1z), sg);
  if (!(den != 0.0f && U >= 0.0f && V >= 0.0f && fadd(U, V) <= ad)) return false;
  const float T = fmul(dot3m(ngx, ngy, ngz, cx, cy, cz), sg);
  return fmul(ad, 0.0f) < T && T <= fmul(ad, r.tfar);
}

}  // namespace c2b
