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

// SnavelyCamera::project, src/baproblem.rs:145-151; |p_|^4 as m2*m2
C2B_HD void project(double f, double k1, double k2, V3 pc, double &u, double &v) {
  double px = ddiv(-pc.x, pc.z), py = ddiv(-pc.y, pc.z);
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

// ---- watertight ray/triangle test (Woop, Benthin, Wald 2013), hit interval 0 < t <= tfar -------
struct Shear {
  int kz;     // dominant axis of the direction
  bool swap;  // dir[kz] < 0: kx and ky trade places (keeps the winding)
  float Sx, Sy, Sz;
};

C2B_HD float sel3(int k, float a, float b, float c) { return k == 0 ? a : (k == 1 ? b : c); }

C2B_HD Shear ray_shear(const Ray &r) {
  Shear s;
  int kz = 0;
  float m = fabsf(r.dx);
  if (fabsf(r.dy) > m) {
    kz = 1;
    m = fabsf(r.dy);
  }
  if (fabsf(r.dz) > m) kz = 2;
  // kx = (kz+1)%3, ky = (kx+1)%3, swapped when the dominant component is negative
  const float dz = sel3(kz, r.dx, r.dy, r.dz);
  float dkx = sel3(kz, r.dy, r.dz, r.dx);
  float dky = sel3(kz, r.dz, r.dx, r.dy);
  s.kz = kz;
  s.swap = dz < 0.0f;
  if (s.swap) {
    const float t = dkx;
    dkx = dky;
    dky = t;
  }
  s.Sx = fdiv(dkx, dz);
  s.Sy = fdiv(dky, dz);
  s.Sz = fdiv(1.0f, dz);
  return s;
}

// components (kx, ky, kz) of a translated vertex, as two rounds of selects (no branches)
C2B_HD void permute(const Shear &s, float a0, float a1, float a2, float &x, float &y, float &z) {
  const bool z0 = s.kz == 0, z1 = s.kz == 1;
  const float rx = z0 ? a1 : (z1 ? a2 : a0);
  const float ry = z0 ? a2 : (z1 ? a0 : a1);
  z = z0 ? a0 : (z1 ? a1 : a2);
  x = s.swap ? ry : rx;
  y = s.swap ? rx : ry;
}

// returns true when the triangle occludes the ray.  t_out (optional) receives T/det for the
// closest-hit entry (c2b_intersect1); it is not part of the occlusion decision.
C2B_HD bool ray_triangle(const Ray &r, const Shear &s, float v0x, float v0y, float v0z, float v1x,
                         float v1y, float v1z, float v2x, float v2y, float v2z,
                         float *t_out = nullptr) {
  float Akx, Aky, Akz, Bkx, Bky, Bkz, Ckx, Cky, Ckz;
  permute(s, fsub(v0x, r.ox), fsub(v0y, r.oy), fsub(v0z, r.oz), Akx, Aky, Akz);
  permute(s, fsub(v1x, r.ox), fsub(v1y, r.oy), fsub(v1z, r.oz), Bkx, Bky, Bkz);
  permute(s, fsub(v2x, r.ox), fsub(v2y, r.oy), fsub(v2z, r.oz), Ckx, Cky, Ckz);
  float Ax = fsub(Akx, fmul(s.Sx, Akz)), Ay = fsub(Aky, fmul(s.Sy, Akz));
  float Bx = fsub(Bkx, fmul(s.Sx, Bkz)), By = fsub(Bky, fmul(s.Sy, Bkz));
  float Cx = fsub(Ckx, fmul(s.Sx, Ckz)), Cy = fsub(Cky, fmul(s.Sy, Ckz));
  float U = fsub(fmul(Cx, By), fmul(Cy, Bx));
  float V = fsub(fmul(Ax, Cy), fmul(Ay, Cx));
  float W = fsub(fmul(Bx, Ay), fmul(By, Ax));
  if (U == 0.0f || V == 0.0f || W == 0.0f) {
    U = d2f(dsub(dmul((double)Cx, (double)By), dmul((double)Cy, (double)Bx)));
    V = d2f(dsub(dmul((double)Ax, (double)Cy), dmul((double)Ay, (double)Cx)));
    W = d2f(dsub(dmul((double)Bx, (double)Ay), dmul((double)By, (double)Ax)));
  }
  if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
  float det = fadd(fadd(U, V), W);
  if (det == 0.0f) return false;
  float Az = fmul(s.Sz, Akz), Bz = fmul(s.Sz, Bkz), Cz = fmul(s.Sz, Ckz);
  float T = fadd(fadd(fmul(U, Az), fmul(V, Bz)), fmul(W, Cz));
  float ad = fabsf(det);
  float Ts = det < 0.0f ? -T : T;
  bool hit = Ts > 0.0f && Ts <= fmul(r.tfar, ad);
  if (hit && t_out) *t_out = fdiv(Ts, ad);
  return hit;
}

}  // namespace c2b
