/* c2b_oracle.c — CPU ORACLE (test infrastructure, see c2b_oracle.h for the rules).
 *
 * Plain C restatement of the city2ba hot path.  Compile with -ffp-contract=off: the
 * reference is Rust, which never contracts a*b+c into an FMA, and the CUDA path
 * uses explicit round-to-nearest intrinsics in the same operation order, so every
 * f64/f32 compare below is bit-reproducible on both sides.
 *
 * PARITY UNPINNED at the Embree boundary (no golden vectors in the reference, Embree
 * 3.8.0 / embree-rs 0.3.6 not available): see header.
 */
#include "c2b_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * small f64 helpers in cgmath 0.17 operation order
 * ---------------------------------------------------------------------------------------- */

/* cgmath Matrix3 * Vector3 = col0*x + col1*y + col2*z, left to right
 * (used by Basis3::rotate_point / rotate_vector, src/baproblem.rs:142,162,168,174) */
static void mat_vec(const double *R, const double *v, double *o) {
  o[0] = (R[0] * v[0] + R[3] * v[1]) + R[6] * v[2];
  o[1] = (R[1] * v[0] + R[4] * v[1]) + R[7] * v[2];
  o[2] = (R[2] * v[0] + R[5] * v[1]) + R[8] * v[2];
}

/* Matrix3 * Matrix3: columns of rhs rotated one by one (src/baproblem.rs:167 `self.dir * delta_dir`) */
static void mat_mat(const double *A, const double *B, double *O) {
  double tmp[9];
  for (int c = 0; c < 3; ++c) mat_vec(A, B + 3 * c, tmp + 3 * c);
  memcpy(O, tmp, sizeof tmp);
}

static void cross3(const double *a, const double *b, double *o) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}

/* cgmath Matrix3::invert (general inverse through cofactors / determinant); Basis3::invert
 * calls it (src/baproblem.rs:162,174).  Output column-major. */
static void mat_invert(const double *M, double *O) {
  const double *c0 = M, *c1 = M + 3, *c2 = M + 6;
  double det = (c0[0] * (c1[1] * c2[2] - c2[1] * c1[2]) - c1[0] * (c0[1] * c2[2] - c2[1] * c0[2])) +
               c2[0] * (c0[1] * c1[2] - c1[1] * c0[2]);
  double r0[3], r1[3], r2[3];
  cross3(c1, c2, r0);
  cross3(c2, c0, r1);
  cross3(c0, c1, r2);
  /* rows of the inverse are r_i/det; store transposed (column-major) */
  for (int i = 0; i < 3; ++i) {
    O[0 + 3 * i] = r0[i] / det;
    O[1 + 3 * i] = r1[i] / det;
    O[2 + 3 * i] = r2[i] / det;
  }
}

static double mag3(const double *v) { return sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]); }

/* cgmath InnerSpace::normalize = self * (1 / magnitude) */
static void normalize3(const double *v, double *o) {
  double s = 1.0 / mag3(v);
  o[0] = v[0] * s;
  o[1] = v[1] * s;
  o[2] = v[2] * s;
}

/* ------------------------------------------------------------------------------------------
 * camera math
 * ---------------------------------------------------------------------------------------- */

/* src/baproblem.rs:141-143 */
void orc_project_world(const double *cam, const double *p, double *o) {
  double r[3];
  mat_vec(cam, p, r);
  o[0] = r[0] + cam[9];
  o[1] = r[1] + cam[10];
  o[2] = r[2] + cam[11];
}

/* src/baproblem.rs:145-151.  `magnitude().powf(4.0)` is restated as m2*m2 (<= 2 ulp from the
 * reference's sqrt(m2)^4, and identical whenever k2 == 0). */
void orc_project(const double *cam, const double *pc, double *uv) {
  double px = (-pc[0]) / pc[2];
  double py = (-pc[1]) / pc[2];
  double m2 = px * px + py * py;
  double r = (1.0 + cam[13] * m2) + cam[14] * (m2 * m2);
  double fr = cam[12] * r;
  uv[0] = fr * px;
  uv[1] = fr * py;
}

/* src/baproblem.rs:161-163 */
void orc_center(const double *cam, double *o) {
  double inv[9], r[3];
  mat_invert(cam, inv);
  mat_vec(inv, cam + 9, r);
  o[0] = -r[0];
  o[1] = -r[1];
  o[2] = -r[2];
}

/* src/baproblem.rs:173-175 */
void orc_to_world(const double *cam, const double *p, double *o) {
  double inv[9], d[3] = {p[0] - cam[9], p[1] - cam[10], p[2] - cam[11]};
  mat_invert(cam, inv);
  mat_vec(inv, d, o);
}

/* src/baproblem.rs:153-159 */
void orc_from_position_direction(const double *pos, const double *R, double *cam) {
  double r[3];
  memmove(cam, R, 9 * sizeof(double));
  mat_vec(cam, pos, r);
  cam[9] = -1.0 * r[0];
  cam[10] = -1.0 * r[1];
  cam[11] = -1.0 * r[2];
  cam[12] = 1.0;
  cam[13] = 0.0;
  cam[14] = 0.0;
}

/* src/baproblem.rs:165-171: dir' = dir*delta ; loc' = -(dir_old * (center + delta_loc)) */
void orc_transform(const double *cam, const double *dR, const double *dloc, double *out) {
  double c[3], q[3], r[3], Rn[9];
  orc_center(cam, c);
  q[0] = c[0] + dloc[0];
  q[1] = c[1] + dloc[1];
  q[2] = c[2] + dloc[2];
  mat_vec(cam, q, r);
  mat_mat(cam, dR, Rn);
  double i0 = cam[12], i1 = cam[13], i2 = cam[14];
  memcpy(out, Rn, sizeof Rn);
  out[9] = -1.0 * r[0];
  out[10] = -1.0 * r[1];
  out[11] = -1.0 * r[2];
  out[12] = i0;
  out[13] = i1;
  out[14] = i2;
}

/* cgmath Matrix3::from_angle_x / from_angle_y / from_axis_angle (column-major) */
void orc_from_angle_x(double rad, double *R) {
  double s = sin(rad), c = cos(rad);
  double m[9] = {1, 0, 0, 0, c, s, 0, -s, c};
  memcpy(R, m, sizeof m);
}
void orc_from_angle_y(double rad, double *R) {
  double s = sin(rad), c = cos(rad);
  double m[9] = {c, 0, -s, 0, 1, 0, s, 0, c};
  memcpy(R, m, sizeof m);
}
void orc_from_axis_angle(const double *a, double rad, double *R) {
  double s = sin(rad), c = cos(rad);
  double k = 1.0 - c;
  double m[9] = {k * a[0] * a[0] + c,        k * a[0] * a[1] + s * a[2], k * a[0] * a[2] - s * a[1],
                 k * a[0] * a[1] - s * a[2], k * a[1] * a[1] + c,        k * a[1] * a[2] + s * a[0],
                 k * a[0] * a[2] + s * a[1], k * a[1] * a[2] - s * a[0], k * a[2] * a[2] + c};
  memcpy(R, m, sizeof m);
}
/* cgmath: Rad::from(Deg(d)) = d * (pi / 180) */
double orc_deg_to_rad(double deg) { return deg * (3.14159265358979323846 / 180.0); }

/* cgmath Quaternion::from(Matrix3) (trace-based, Shepperd branches); returns (s, x, y, z) */
static void quat_from_mat(const double *m, double *q) {
  /* m[c*3+r] column-major: mat[c][r] */
#define MAT(c, r) m[(c)*3 + (r)]
  double trace = (MAT(0, 0) + MAT(1, 1)) + MAT(2, 2);
  if (trace >= 0.0) {
    double s = sqrt(1.0 + trace);
    double w = 0.5 * s;
    s = 0.5 / s;
    q[0] = w;
    q[1] = (MAT(1, 2) - MAT(2, 1)) * s;
    q[2] = (MAT(2, 0) - MAT(0, 2)) * s;
    q[3] = (MAT(0, 1) - MAT(1, 0)) * s;
  } else if (MAT(0, 0) > MAT(1, 1) && MAT(0, 0) > MAT(2, 2)) {
    double s = sqrt(((MAT(0, 0) - MAT(1, 1)) - MAT(2, 2)) + 1.0);
    double x = 0.5 * s;
    s = 0.5 / s;
    q[1] = x;
    q[2] = (MAT(1, 0) + MAT(0, 1)) * s;
    q[3] = (MAT(0, 2) + MAT(2, 0)) * s;
    q[0] = (MAT(1, 2) - MAT(2, 1)) * s;
  } else if (MAT(1, 1) > MAT(2, 2)) {
    double s = sqrt(((MAT(1, 1) - MAT(0, 0)) - MAT(2, 2)) + 1.0);
    double y = 0.5 * s;
    s = 0.5 / s;
    q[2] = y;
    q[3] = (MAT(2, 1) + MAT(1, 2)) * s;
    q[1] = (MAT(1, 0) + MAT(0, 1)) * s;
    q[0] = (MAT(2, 0) - MAT(0, 2)) * s;
  } else {
    double s = sqrt(((MAT(2, 2) - MAT(0, 0)) - MAT(1, 1)) + 1.0);
    double z = 0.5 * s;
    s = 0.5 / s;
    q[3] = z;
    q[1] = (MAT(0, 2) + MAT(2, 0)) * s;
    q[2] = (MAT(2, 1) + MAT(1, 2)) * s;
    q[0] = (MAT(0, 1) - MAT(1, 0)) * s;
  }
#undef MAT
}

/* cgmath Matrix3::from(Quaternion) (column-major) */
static void mat_from_quat(const double *q, double *R) {
  double s = q[0], x = q[1], y = q[2], z = q[3];
  double x2 = x + x, y2 = y + y, z2 = z + z;
  double xx2 = x2 * x, xy2 = x2 * y, xz2 = x2 * z;
  double yy2 = y2 * y, yz2 = y2 * z, zz2 = z2 * z;
  double sy2 = y2 * s, sz2 = z2 * s, sx2 = x2 * s;
  double m[9] = {1.0 - yy2 - zz2, xy2 + sz2,       xz2 - sy2,      xy2 - sz2, 1.0 - xx2 - zz2,
                 yz2 + sx2,       xz2 + sy2,       yz2 - sx2,      1.0 - xx2 - yy2};
  memcpy(R, m, sizeof m);
}

/* src/baproblem.rs:78-90 */
void orc_from_rodrigues(const double *x, double *R) {
  double theta2 = (x[0] * x[0] + x[1] * x[1]) + x[2] * x[2];
  if (theta2 > 2.220446049250313e-16) { /* Rad::<f64>::default_epsilon() == f64::EPSILON */
    double angle = mag3(x);
    double axis[3];
    normalize3(x, axis);
    orc_from_axis_angle(axis, angle, R);
  } else {
    /* Matrix3::new(1, x2, -x1, -x2, 1, x0, x1, -x0, 1) column-major, through a quaternion */
    double m[9] = {1.0, x[2], -x[1], -x[2], 1.0, x[0], x[1], -x[0], 1.0};
    double q[4];
    quat_from_mat(m, q);
    mat_from_quat(q, R);
  }
}

/* src/baproblem.rs:93-102 */
void orc_to_rodrigues(const double *R, double *v) {
  double q[4];
  quat_from_mat(R, q);
  double angle = 2.0 * acos(q[0]);
  double d = 1.0 - q[0] * q[0];
  if (d < 2.220446049250313e-16) {
    v[0] = v[1] = v[2] = 0.0;
  } else {
    double sd = sqrt(d);
    double axis[3] = {q[1] / sd, q[2] / sd, q[3] / sd}, n[3];
    normalize3(axis, n);
    v[0] = n[0] * angle;
    v[1] = n[1] * angle;
    v[2] = n[2] * angle;
  }
}

/* src/baproblem.rs:180-186 */
void orc_from_vec(const double *x, double *cam) {
  orc_from_rodrigues(x, cam);
  for (int i = 0; i < 6; ++i) cam[9 + i] = x[3 + i];
}
/* src/baproblem.rs:189-202 */
void orc_to_vec(const double *cam, double *x) {
  orc_to_rodrigues(cam, x);
  for (int i = 0; i < 6; ++i) x[3 + i] = cam[9 + i];
}

/* ------------------------------------------------------------------------------------------
 * ray construction + watertight ray/triangle predicate (f32, fixed order, explicit fmaf only)
 * ---------------------------------------------------------------------------------------- */

/* src/generate.rs:456-464 */
void orc_make_ray(const double *c, const double *p, orc_ray *ray) {
  double d[3] = {p[0] - c[0], p[1] - c[1], p[2] - c[2]};
  double n = mag3(d);
  double s = 1.0 / n; /* dir.normalize() */
  ray->org[0] = (float)c[0];
  ray->org[1] = (float)c[1];
  ray->org[2] = (float)c[2];
  ray->dir[0] = (float)(d[0] * s);
  ray->dir[1] = (float)(d[1] * s);
  ray->dir[2] = (float)(d[2] * s);
  ray->tfar = (float)n - 1e-6f;
}

/* Watertight ray/triangle predicate, hit interval 0 < t <= tfar (Embree: |den|*tnear < T <=
 * |den|*tfar with tnear = 0), no backface culling.  Edge functions in the form of Embree 3's robust
 * ("Pluecker") triangle intersector — the reference links Embree 3.8.0 through embree-rs 0.3.6
 * (Cargo.lock:278-279), whose source is not vendored, so this is a restatement of the published
 * algorithm, not of the binary (PARITY UNPINNED at this boundary, see the header):
 *   A = v0-o, B = v1-o, C = v2-o;  e0 = C-A, e1 = A-B, e2 = B-C
 *   U = d.(e0 x (C+A)),  V = d.(e1 x (A+B)),  W = d.(e2 x (B+C));  det = U+V+W;  T = 2 A.(e0 x e1)
 * Cross products unfused (exactly antisymmetric), dot products as fmaf chains: the record of a
 * triangle depends on (origin, triangle) only, which the CUDA path exploits per camera.
 * Fixed operation order = c2b_math.cuh tri_record / ray_tri_record. */
typedef struct {
  float u[3], v[3], w[3], T;
} tri_rec;

static float dot3f(const float *a, const float *b) { return fmaf(a[2], b[2], fmaf(a[1], b[1], a[0] * b[0])); }

static void cross3f(const float *a, const float *b, float *o) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}

static void tri_record(const float *org, const float *v0, const float *v1, const float *v2, tri_rec *t) {
  float A[3], B[3], C[3], e0[3], e1[3], e2[3], su[3], sv[3], sw[3], n[3];
  for (int k = 0; k < 3; ++k) {
    A[k] = v0[k] - org[k];
    B[k] = v1[k] - org[k];
    C[k] = v2[k] - org[k];
  }
  for (int k = 0; k < 3; ++k) {
    e0[k] = C[k] - A[k];
    e1[k] = A[k] - B[k];
    e2[k] = B[k] - C[k];
    su[k] = C[k] + A[k];
    sv[k] = A[k] + B[k];
    sw[k] = B[k] + C[k];
  }
  cross3f(e0, su, t->u);
  cross3f(e1, sv, t->v);
  cross3f(e2, sw, t->w);
  cross3f(e0, e1, n);
  t->T = 2.0f * dot3f(A, n);
}

/* Outputs the scaled quantities for callers that want them. */
static int tri_test(const orc_ray *r, const float *v0, const float *v1, const float *v2, float *oU,
                    float *oV, float *oW, float *oT) {
  tri_rec t;
  tri_record(r->org, v0, v1, v2, &t);
  float U = dot3f(r->dir, t.u), V = dot3f(r->dir, t.v), W = dot3f(r->dir, t.w);
  if (oU) {
    *oU = U;
    *oV = V;
    *oW = W;
    *oT = t.T;
  }
  if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return 0;
  float det = (U + V) + W;
  if (det == 0.0f) return 0;
  float ad = fabsf(det);
  float Ts = det < 0.0f ? -t.T : t.T;
  return (Ts > 0.0f && Ts <= r->tfar * ad) ? 1 : 0; /* NaN anywhere => 0 (ray stays visible) */
}

int orc_ray_triangle(const orc_ray *ray, const float *v0, const float *v1, const float *v2) {
  return tri_test(ray, v0, v1, v2, 0, 0, 0, 0);
}

/* flagged-epsilon classification of one (ray, triangle) pair, all in f64 on the f32 inputs */
#define ORC_EPS 1e-5
static unsigned tri_flags(const orc_ray *r, const float *v0, const float *v1, const float *v2) {
  double o[3] = {r->org[0], r->org[1], r->org[2]}, d[3] = {r->dir[0], r->dir[1], r->dir[2]};
  double a[3] = {v0[0], v0[1], v0[2]}, b[3] = {v1[0], v1[1], v1[2]}, c[3] = {v2[0], v2[1], v2[2]};
  double e1[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]};
  double e2[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
  double n[3];
  cross3(e1, e2, n);
  double nn = mag3(n), dn = mag3(d);
  if (!(nn > 0.0) || !(dn > 0.0)) return 0;
  double tf = (double)r->tfar;
  double cosang = ((n[0] * d[0] + n[1] * d[1]) + n[2] * d[2]) / (nn * dn);
  double oa[3] = {a[0] - o[0], a[1] - o[1], a[2] - o[2]};
  double pd = ((n[0] * oa[0] + n[1] * oa[1]) + n[2] * oa[2]) / nn; /* signed plane distance */
  unsigned f = 0;
  if (fabs(cosang) <= ORC_EPS) {
    /* grazing: nearly parallel and the segment runs within eps of the plane and overlaps the
     * triangle's (inflated) bounding box */
    double scale = 1.0 + tf;
    if (fabs(pd) <= ORC_EPS * scale) {
      int overlap = 1;
      for (int k = 0; k < 3; ++k) {
        double lo = fmin(a[k], fmin(b[k], c[k])) - ORC_EPS * scale;
        double hi = fmax(a[k], fmax(b[k], c[k])) + ORC_EPS * scale;
        double s0 = o[k], s1 = o[k] + d[k] * tf;
        if (fmax(s0, s1) < lo || fmin(s0, s1) > hi) overlap = 0;
      }
      if (overlap) f |= ORC_FLAG_GRAZE;
    }
    return f;
  }
  double t = pd / (cosang * dn); /* distance along the (unnormalised-safe) ray */
  double h[3] = {o[0] + d[0] * t - a[0], o[1] + d[1] * t - a[1], o[2] + d[2] * t - a[2]};
  /* barycentrics of the hit point */
  double d11 = (e1[0] * e1[0] + e1[1] * e1[1]) + e1[2] * e1[2];
  double d12 = (e1[0] * e2[0] + e1[1] * e2[1]) + e1[2] * e2[2];
  double d22 = (e2[0] * e2[0] + e2[1] * e2[1]) + e2[2] * e2[2];
  double h1 = (h[0] * e1[0] + h[1] * e1[1]) + h[2] * e1[2];
  double h2 = (h[0] * e2[0] + h[1] * e2[1]) + h[2] * e2[2];
  double den = d11 * d22 - d12 * d12;
  if (!(den != 0.0)) return 0;
  double bv = (d22 * h1 - d12 * h2) / den, bw = (d11 * h2 - d12 * h1) / den, bu = 1.0 - bv - bw;
  double mn = fmin(bu, fmin(bv, bw));
  /* conditioning: t = (N.C)/(N.D) and the barycentrics lose a factor 1/|cos| of the f32 precision
   * when the ray runs nearly along the face, so the margins widen with it: 1e-5 down to
   * |cos| = 0.05, 5e-7/|cos| below (found by the Moeller-Trumbore A/B count: a ray at
   * |cos| = 2.7e-4 whose t differed from tfar by 9e-6 relative was decided differently by the two
   * f32 predicates while outside the fixed 1e-5 margin) */
  double eps = ORC_EPS * fmax(1.0, 0.05 / fabs(cosang));
  if (mn < -eps) return 0; /* clearly outside */
  int t_in = (t > -eps * (1.0 + tf)) && (t <= tf * (1.0 + eps));
  if (mn <= eps && t_in) f |= ORC_FLAG_EDGE;
  if (fabs(t - tf) <= eps * fmax(1.0, tf)) f |= ORC_FLAG_ENDPOINT;
  return f;
}

/* ------------------------------------------------------------------------------------------
 * triangle soup helpers
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  float v[9];
} tri9;

/* src/generate.rs:74-105: f32 positions, u32 index triples.  Triples with a repeated index
 * (e.g. the 2-index `l` records tobj leaves in a path object) are zero-area and can never
 * be hit; they are dropped here, like c2b_scene_create does. */
static tri9 *gather_tris(const float *xyz, uint64_t nv, const uint32_t *tri, uint64_t nt,
                         uint64_t *n_out) {
  tri9 *t = (tri9 *)malloc(sizeof(tri9) * (nt ? nt : 1));
  uint64_t n = 0;
  for (uint64_t i = 0; i < nt; ++i) {
    uint32_t a = tri[3 * i], b = tri[3 * i + 1], c = tri[3 * i + 2];
    if (a == b || b == c || a == c) continue;
    if (a >= nv || b >= nv || c >= nv) continue;
    memcpy(t[n].v + 0, xyz + 3 * (uint64_t)a, 12);
    memcpy(t[n].v + 3, xyz + 3 * (uint64_t)b, 12);
    memcpy(t[n].v + 6, xyz + 3 * (uint64_t)c, 12);
    ++n;
  }
  *n_out = n;
  return t;
}

/* ------------------------------------------------------------------------------------------
 * cull + projection predicates shared by generate.rs:450-454 and synthetic.rs:287-289
 * ---------------------------------------------------------------------------------------- */
static int cull_pair(const double *cam, const double *center, const double *p, double max_dist,
                     double *uv, int *near_boundary) {
  double pc[3];
  orc_project_world(cam, p, pc);
  double d[3] = {center[0] - p[0], center[1] - p[1], center[2] - p[2]};
  double dist = mag3(d);
  if (near_boundary) {
    *near_boundary = 0;
    if (fabs(dist - max_dist) <= 1e-12 * max_dist) *near_boundary = 1;
  }
  if (!(dist < max_dist && pc[2] <= 0.0)) {
    if (near_boundary && dist < max_dist * (1.0 + 1e-12) && fabs(pc[2]) <= 1e-12 * (1.0 + dist))
      *near_boundary = 1;
    return 0;
  }
  orc_project(cam, pc, uv);
  if (near_boundary) {
    if (fabs(pc[2]) <= 1e-12 * (1.0 + dist)) *near_boundary = 1;
    if (fabs(fabs(uv[0]) - 1.0) <= 1e-12 || fabs(fabs(uv[1]) - 1.0) <= 1e-12) *near_boundary = 1;
  }
  return (uv[0] >= -1.0 && uv[0] <= 1.0 && uv[1] >= -1.0 && uv[1] <= 1.0) ? 1 : 0;
}

/* growable arrays */
typedef struct {
  uint64_t *pt;
  double *uv;
  uint8_t *occ, *flg;
  uint64_t n, cap;
} candvec;
static void cv_push(candvec *v, uint64_t pt, const double *uv, uint8_t occ, uint8_t flg) {
  if (v->n == v->cap) {
    v->cap = v->cap ? v->cap * 2 : 1024;
    v->pt = (uint64_t *)realloc(v->pt, v->cap * 8);
    v->uv = (double *)realloc(v->uv, v->cap * 16);
    v->occ = (uint8_t *)realloc(v->occ, v->cap);
    v->flg = (uint8_t *)realloc(v->flg, v->cap);
  }
  v->pt[v->n] = pt;
  v->uv[2 * v->n] = uv[0];
  v->uv[2 * v->n + 1] = uv[1];
  v->occ[v->n] = occ;
  v->flg[v->n] = flg;
  v->n++;
}

static void finish_visible(orc_vis *r, uint64_t C) {
  r->offsets = (uint64_t *)calloc(C + 1, 8);
  uint64_t nobs = 0;
  for (uint64_t i = 0; i < r->n_candidates; ++i) nobs += !r->cand_occluded[i];
  r->n_obs = nobs;
  r->point_idx = (uint64_t *)malloc(8 * (nobs ? nobs : 1));
  r->uv = (double *)malloc(16 * (nobs ? nobs : 1));
  uint64_t o = 0;
  for (uint64_t c = 0; c < C; ++c) {
    r->offsets[c] = o;
    for (uint64_t i = r->cand_offsets[c]; i < r->cand_offsets[c + 1]; ++i) {
      if (r->cand_occluded[i]) continue;
      r->point_idx[o] = r->cand_point[i];
      r->uv[2 * o] = r->cand_uv[2 * i];
      r->uv[2 * o + 1] = r->cand_uv[2 * i + 1];
      ++o;
    }
  }
  r->offsets[C] = o;
}

/* src/generate.rs:424-481, one camera after the other, brute force over all triangles */
orc_vis *orc_visibility_graph(const float *xyz, uint64_t nv, const uint32_t *tri, uint64_t nt,
                              const double *cams, uint64_t C, const double *pts, uint64_t P,
                              double max_dist, int endpoint_guard_rel, int want_flags) {
  uint64_t ntri = 0;
  tri9 *T = gather_tris(xyz, nv, tri, nt, &ntri);
  orc_vis *r = (orc_vis *)calloc(1, sizeof(orc_vis));
  r->n_cameras = C;
  r->cand_offsets = (uint64_t *)calloc(C + 1, 8);
  candvec cv = {0};
  for (uint64_t c = 0; c < C; ++c) {
    const double *cam = cams + ORC_CAM_STRIDE * c;
    double center[3];
    orc_center(cam, center); /* the reference recomputes it per pair; same value */
    r->cand_offsets[c] = cv.n;
    for (uint64_t i = 0; i < P; ++i) {
      const double *p = pts + 3 * i;
      double uv[2];
      int nb = 0;
      int keep = cull_pair(cam, center, p, max_dist, uv, want_flags ? &nb : 0);
      if (want_flags && nb) r->n_flag_cull++;
      if (!keep) continue;
      orc_ray ray;
      orc_make_ray(center, p, &ray);
      if (endpoint_guard_rel) ray.tfar = ray.tfar * (1.0f - 3.814697265625e-06f); /* 2^-18 */
      uint8_t occ = 0, flg = (want_flags && nb) ? ORC_FLAG_CULL : 0;
      for (uint64_t t = 0; t < ntri; ++t) {
        if (tri_test(&ray, T[t].v, T[t].v + 3, T[t].v + 6, 0, 0, 0, 0)) {
          occ = 1;
          if (!want_flags) break;
        }
        if (want_flags) flg |= (uint8_t)tri_flags(&ray, T[t].v, T[t].v + 3, T[t].v + 6);
      }
      if (flg & ORC_FLAG_EDGE) r->n_flag_edge++;
      if (flg & ORC_FLAG_GRAZE) r->n_flag_graze++;
      if (flg & ORC_FLAG_ENDPOINT) r->n_flag_endpoint++;
      cv_push(&cv, i, uv, occ, flg);
    }
  }
  r->cand_offsets[C] = cv.n;
  r->n_candidates = cv.n;
  r->cand_point = cv.pt ? cv.pt : (uint64_t *)malloc(8);
  r->cand_uv = cv.uv ? cv.uv : (double *)malloc(16);
  r->cand_occluded = cv.occ ? cv.occ : (uint8_t *)malloc(1);
  r->cand_flags = cv.flg ? cv.flg : (uint8_t *)malloc(1);
  finish_visible(r, C);
  free(T);
  return r;
}

/* ---- A/B counter against Embree's DEFAULT triangle intersector -------------------------------------
 * The reference builds a default (non-robust) scene, for which Embree 3 tests triangles with its
 * Moeller-Trumbore intersector, not the watertight Pluecker form used above.  Embree 3.8.0 is not
 * vendored, so this is a restatement of the published algorithm (embree3 kernels/geometry
 * triangle_intersector_moeller.h, from memory — UNPINNED like everything at that boundary): with the
 * stored v0, e1 = v0 - v1, e2 = v2 - v0, Ng = e2 x e1,
 *     C = v0 - org, R = C x dir, den = Ng . dir, U = (R . e2) ^ sign(den), V = (R . e1) ^ sign(den),
 *     hit  <=>  den != 0  and  U >= 0  and  V >= 0  and  U + V <= |den|
 *               and  |den| * tnear < T <= |den| * tfar   with  T = (Ng . C) ^ sign(den), tnear = 0,
 * all in f32, dot products as Embree's fused madd chains.  It exists to COUNT the rays on which the two
 * predicates disagree and to check that every such ray carries a flag (tests/test_oracle_visibility.py). */
static float dot3m(const float *a, const float *b) { return fmaf(a[0], b[0], fmaf(a[1], b[1], a[2] * b[2])); }
int orc_ray_triangle_mt(const orc_ray *r, const float *v0, const float *v1, const float *v2) {
  float e1[3], e2[3], Ng[3], Cv[3], R[3];
  for (int k = 0; k < 3; ++k) {
    e1[k] = v0[k] - v1[k];
    e2[k] = v2[k] - v0[k];
    Cv[k] = v0[k] - r->org[k];
  }
  cross3f(e2, e1, Ng);
  cross3f(Cv, r->dir, R);
  const float den = dot3m(Ng, r->dir);
  const float ad = fabsf(den);
  const float sg = den < 0.0f ? -1.0f : 1.0f; /* xor with the sign bit of den */
  const float U = dot3m(R, e2) * sg, V = dot3m(R, e1) * sg;
  if (!(den != 0.0f && U >= 0.0f && V >= 0.0f && U + V <= ad)) return 0;
  const float T = dot3m(Ng, Cv) * sg;
  return (ad * 0.0f < T && T <= ad * r->tfar) ? 1 : 0;
}

/* occluded[] (one byte per candidate of v, same order) under the Moeller-Trumbore restatement */
void orc_occluded_mt(const float *xyz, uint64_t nv, const uint32_t *tri, uint64_t nt, const double *cams,
                     uint64_t C, const double *pts, const uint64_t *cand_offsets, const uint64_t *cand_point,
                     int endpoint_guard_rel, uint8_t *occluded) {
  uint64_t ntri = 0;
  tri9 *T = gather_tris(xyz, nv, tri, nt, &ntri);
  for (uint64_t c = 0; c < C; ++c) {
    double center[3];
    orc_center(cams + ORC_CAM_STRIDE * c, center);
    for (uint64_t k = cand_offsets[c]; k < cand_offsets[c + 1]; ++k) {
      orc_ray ray;
      orc_make_ray(center, pts + 3 * cand_point[k], &ray);
      if (endpoint_guard_rel) ray.tfar = ray.tfar * (1.0f - 3.814697265625e-06f);
      uint8_t occ = 0;
      for (uint64_t t = 0; t < ntri && !occ; ++t) occ = (uint8_t)orc_ray_triangle_mt(&ray, T[t].v, T[t].v + 3, T[t].v + 6);
      occluded[k] = occ;
    }
  }
  free(T);
}

void orc_vis_free(orc_vis *v) {
  if (!v) return;
  free(v->cand_offsets);
  free(v->cand_point);
  free(v->cand_uv);
  free(v->cand_occluded);
  free(v->cand_flags);
  free(v->offsets);
  free(v->point_idx);
  free(v->uv);
  free(v);
}

/* ------------------------------------------------------------------------------------------
 * synthetic lattice (src/synthetic.rs)
 * ---------------------------------------------------------------------------------------- */
uint64_t orc_grid_num_cameras(uint64_t cpb, uint64_t n) { return 4 * cpb * n * (n + 1); }
uint64_t orc_grid_num_points(uint64_t ppb, uint64_t n) { return 12 * ppb * n * (n + 1); }

/* src/synthetic.rs:178-210 */
void orc_grid_cameras(uint64_t cpb, uint64_t n, double L, double h, double *out) {
  double Rm90[9], Rp90[9], R180[9], R1[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  orc_from_angle_y(orc_deg_to_rad(-90.), Rm90);
  orc_from_angle_y(orc_deg_to_rad(90.), Rp90);
  orc_from_angle_y(orc_deg_to_rad(180.), R180);
  uint64_t k = 0;
  for (uint64_t bx = 0; bx <= n; ++bx) {
    double ox = L * (double)bx;
    for (uint64_t by = 0; by <= n; ++by) {
      double oz = L * (double)by;
      for (uint64_t i = 0; i < cpb; ++i) {
        if (bx != n) {
          double pos[3] = {ox + (double)i / (double)cpb * L, h, oz};
          orc_from_position_direction(pos, Rm90, out + ORC_CAM_STRIDE * k++);
          orc_from_position_direction(pos, Rp90, out + ORC_CAM_STRIDE * k++);
        }
        if (by != n) {
          double pos[3] = {ox, h, oz + (double)i / (double)cpb * L};
          orc_from_position_direction(pos, R180, out + ORC_CAM_STRIDE * k++);
          orc_from_position_direction(pos, R1, out + ORC_CAM_STRIDE * k++);
        }
      }
    }
  }
}

/* src/synthetic.rs:213-258 */
void orc_grid_points(uint64_t ppb, uint64_t n, double L, double inset, double ph, double *out) {
  uint64_t k = 0;
#define PUSH(X, Y, Z) \
  do {                \
    out[3 * k] = (X); \
    out[3 * k + 1] = (Y); \
    out[3 * k + 2] = (Z); \
    ++k;              \
  } while (0)
  for (uint64_t bx = 0; bx <= n; ++bx) {
    double ox = L * (double)bx;
    for (uint64_t by = 0; by <= n; ++by) {
      double oz = L * (double)by;
      for (uint64_t i = 0; i < ppb; ++i) {
        double step = (L - inset * 2.) / (double)ppb;
        if (bx != n) {
          double lx = ox + inset + (double)i * step;
          PUSH(lx, ph, oz - inset);
          PUSH(lx, ph, oz + inset);
          PUSH(lx + step / 2., 0., oz - inset);
          PUSH(lx + step / 2., 0., oz + inset);
          PUSH(lx + step / 2., 0., oz - inset / 2.);
          PUSH(lx + step / 2., 0., oz + inset / 2.);
        }
        if (by != n) {
          double lz = oz + inset + (double)i * step;
          PUSH(ox - inset, ph, lz);
          PUSH(ox + inset, ph, lz);
          PUSH(ox - inset, 0., lz + step / 2.);
          PUSH(ox + inset, 0., lz + step / 2.);
          PUSH(ox - inset / 2., 0., lz + step / 2.);
          PUSH(ox + inset / 2., 0., lz + step / 2.);
        }
      }
    }
  }
#undef PUSH
}

/* src/synthetic.rs:323-333 */
void orc_line_cameras(uint64_t nc, double length, double h, double *out) {
  double R180[9];
  orc_from_angle_y(orc_deg_to_rad(180.), R180);
  for (uint64_t i = 0; i < nc; ++i) {
    double pos[3] = {0., h, (double)i * length / (double)(nc - 1)};
    orc_from_position_direction(pos, R180, out + ORC_CAM_STRIDE * i);
  }
}
/* src/synthetic.rs:334-344 */
void orc_line_points(uint64_t np, double length, double off, double ph, double *out) {
  for (uint64_t i = 0; i < np; ++i) {
    double z = (double)(i / 2) * length / (double)(np / 2 - 1);
    out[3 * i] = (i % 2 == 0) ? -off : off;
    out[3 * i + 1] = ph;
    out[3 * i + 2] = z;
  }
}

/* line_intersection 0.4.0 `LineInterval::line_segment(a).relate(&line_segment(b))
 * .unique_intersection()` restated from its published algorithm: solve a.start + t*da =
 * b.start + u*db; parallel (zero cross product) => no unique intersection; a point is returned
 * iff t and u both lie in the closed interval [0,1]; the point is a.start + t*da. */
static int seg_intersect(double ax, double ay, double bx, double by, double cx, double cy,
                         double dx, double dy, double *px, double *py) {
  double dax = bx - ax, day = by - ay, dbx = dx - cx, dby = dy - cy;
  double denom = dax * dby - day * dbx; /* cross(da, db) */
  if (denom == 0.0) return 0;
  double sx = cx - ax, sy = cy - ay;
  /* t = cross(q-p, s / rxs), u = cross(q-p, r / rxs): the crate divides the direction first */
  double t = sx * (dby / denom) - sy * (dbx / denom);
  double u = sx * (day / denom) - sy * (dax / denom);
  if (t < 0.0 || t > 1.0 || u < 0.0 || u > 1.0) return 0;
  *px = ax + t * dax;
  *py = ay + t * day;
  return 1;
}

/* src/synthetic.rs:52-98 (incl. the un-squared second term at :93: sqrt of a negative number is
 * NaN, and NaN > 1e-8 is false) */
static int hits_in_block(double sx, double sy, double ex, double ey, long bxi, long byi, double L,
                         double inset) {
  double be = L - inset, ox = (double)bxi * L, oy = (double)byi * L;
  double sides[4][4] = {{ox + inset, oy + inset, ox + inset, oy + be},
                        {ox + inset, oy + inset, ox + be, oy + inset},
                        {ox + be, oy + inset, ox + be, oy + be},
                        {ox + inset, oy + be, ox + be, oy + be}};
  int any = 0;
  for (int k = 0; k < 4; ++k) {
    double px, py;
    if (seg_intersect(sx, sy, ex, ey, sides[k][0], sides[k][1], sides[k][2], sides[k][3], &px, &py)) {
      double dx = ex - px;
      double v = sqrt(dx * dx + (ey - py));
      if (v > 1e-8) any = 1;
    }
  }
  return any;
}

/* src/synthetic.rs:100-124 */
int orc_hits_building(const double *c, const double *p, double L, double inset) {
  double sx = c[0], sy = c[2], ex = p[0], ey = p[2];
  long cbx = (long)trunc(sx / L), cby = (long)trunc(sy / L);
  long pbx = (long)trunc(ex / L), pby = (long)trunc(ey / L);
  long x0 = cbx < pbx ? cbx : pbx, x1 = cbx < pbx ? pbx : cbx;
  long y0 = cby < pby ? cby : pby, y1 = cby < pby ? pby : cby;
  for (long x = x0; x <= x1; ++x)
    for (long y = y0; y <= y1; ++y)
      if (hits_in_block(sx, sy, ex, ey, x, y, L, inset)) return 1;
  return 0;
}

/* src/synthetic.rs:268-297 / 353-379.  The R-tree radius query is a pure pre-filter of the
 * `< max_dist` predicate (squared distance <= max_dist^2 is implied by it), so a scan over all
 * points gives the same set. */
orc_vis *orc_synthetic_visibility(const double *cams, uint64_t C, const double *pts, uint64_t P,
                                  double max_dist, int analytic, double L, double inset) {
  orc_vis *r = (orc_vis *)calloc(1, sizeof(orc_vis));
  r->n_cameras = C;
  r->cand_offsets = (uint64_t *)calloc(C + 1, 8);
  candvec cv = {0};
  for (uint64_t c = 0; c < C; ++c) {
    const double *cam = cams + ORC_CAM_STRIDE * c;
    double center[3];
    orc_center(cam, center);
    r->cand_offsets[c] = cv.n;
    for (uint64_t i = 0; i < P; ++i) {
      const double *p = pts + 3 * i;
      double uv[2];
      if (!cull_pair(cam, center, p, max_dist, uv, 0)) continue;
      uint8_t occ = analytic ? (uint8_t)orc_hits_building(center, p, L, inset) : 0;
      cv_push(&cv, i, uv, occ, 0);
    }
  }
  r->cand_offsets[C] = cv.n;
  r->n_candidates = cv.n;
  r->cand_point = cv.pt ? cv.pt : (uint64_t *)malloc(8);
  r->cand_uv = cv.uv ? cv.uv : (double *)malloc(16);
  r->cand_occluded = cv.occ ? cv.occ : (uint8_t *)malloc(1);
  r->cand_flags = cv.flg ? cv.flg : (uint8_t *)malloc(1);
  finish_visible(r, C);
  return r;
}

/* ------------------------------------------------------------------------------------------
 * city-block box mesh (no counterpart in the reference)
 * ---------------------------------------------------------------------------------------- */
void orc_city_mesh(uint64_t n, double L, double inset, double H, float *xyz, uint32_t *tri) {
  /* vertex k of a box: bit0 -> x hi, bit1 -> y hi, bit2 -> z hi */
  static const uint32_t F[12][3] = {{0, 2, 1}, {1, 2, 3},  /* z lo wall */
                                    {4, 5, 6}, {5, 7, 6},  /* z hi wall */
                                    {0, 4, 2}, {2, 4, 6},  /* x lo wall */
                                    {1, 3, 5}, {3, 7, 5},  /* x hi wall */
                                    {0, 1, 4}, {1, 5, 4},  /* floor */
                                    {2, 6, 3}, {3, 6, 7}}; /* roof */
  uint64_t b = 0;
  for (uint64_t bx = 0; bx < n; ++bx)
    for (uint64_t bz = 0; bz < n; ++bz, ++b) {
      double x0 = (double)bx * L + inset, x1 = (double)(bx + 1) * L - inset;
      double z0 = (double)bz * L + inset, z1 = (double)(bz + 1) * L - inset;
      for (uint32_t k = 0; k < 8; ++k) {
        xyz[3 * (8 * b + k) + 0] = (float)((k & 1) ? x1 : x0);
        xyz[3 * (8 * b + k) + 1] = (float)((k & 2) ? H : 0.0);
        xyz[3 * (8 * b + k) + 2] = (float)((k & 4) ? z1 : z0);
      }
      for (uint32_t f = 0; f < 12; ++f)
        for (int j = 0; j < 3; ++j) tri[3 * (12 * b + f) + j] = (uint32_t)(8 * b) + F[f][j];
    }
}

/* ------------------------------------------------------------------------------------------
 * noise
 * ---------------------------------------------------------------------------------------- */
/* Philox4x32-10 (Salmon et al., SC'11), the published constants */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0;
    c1 = n1;
    c2 = n2;
    c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0;
  out[1] = c1;
  out[2] = c2;
  out[3] = c3;
}

/* ---- the noise draws -------------------------------------------------------------------------
 * The reference draws a direction as a NORMALISED Gaussian pair / triple (unit_random, src/noise.rs:35-43;
 * (nx, ny) / |(nx, ny)|, :159-163) and a magnitude as Normal(mean, std) (rand 0.6.5 Ziggurat), all from the
 * unseedable thread_rng(): only distributions can be matched.  A normalised Gaussian pair IS a uniform
 * direction on the circle, a normalised triple a uniform point on the sphere, so one Philox4x32-10 block
 * per element supplies
 *     circle   (cos, sin)(2 pi w / 2^32)                                  from one 32-bit word,
 *     sphere   z = 1 - 2 (b + 1/2) / 2^32 (Archimedes), azimuth from a    from two words,
 *     N(0,1)   Box-Muller: sqrt(-2 ln u1) cos(theta), u1 = (n40 + 1) 2^-40 in (0, 1], theta from 24 bits,
 *                                                                         from two words,
 * evaluated by table + short polynomial (cos / sin of k 2pi/256 and a degree-5/6 Taylor remainder; ln by
 * 128 intervals of the mantissa and a degree-6 series) with every operation an explicit fma / mul / add in
 * a fixed order.  The CUDA path (c2b_noise.cuh) performs the same operations on the same tables, so points
 * and observations agree with this file bit for bit; the functions themselves are pinned against libm in
 * tests/test_oracle_noise.py (<= 8e-16 absolute) and the distributions by moment / KS tests. */
static double SC_TAB[256][2]; /* cos, sin of k * 2 pi / 256 */
static double LN_TAB[128][2]; /* 1 / c_j, ln c_j with c_j = 1 + (2 j + 1) / 256 */
static int noise_tabs_ready = 0;

void orc_noise_tables(double *sc, double *ln) {
  if (!noise_tabs_ready) {
    for (int k = 0; k < 256; ++k) {
      const double a = (double)k * (6.283185307179586 / 256.0);
      SC_TAB[k][0] = cos(a);
      SC_TAB[k][1] = sin(a);
    }
    for (int j = 0; j < 128; ++j) {
      const double c = 1.0 + (double)(2 * j + 1) / 256.0;
      LN_TAB[j][0] = 1.0 / c;
      LN_TAB[j][1] = log(c);
    }
    noise_tabs_ready = 1;
  }
  if (sc) memcpy(sc, SC_TAB, sizeof SC_TAB);
  if (ln) memcpy(ln, LN_TAB, sizeof LN_TAB);
}

/* (cos, sin) of 2 pi w / 2^32 */
void orc_unit2(uint32_t w, double *c, double *s) {
  if (!noise_tabs_ready) orc_noise_tables(NULL, NULL);
  const uint32_t k = (w + 0x800000u) >> 24; /* nearest table angle; wraps to 0 at the top */
  const double d = (double)(int32_t)(w - (k << 24)) * 1.4629180792671596e-09; /* 2 pi / 2^32; |d| <= pi / 256 */
  const double d2 = d * d;
  double ps = fma(d2, 1.0 / 120.0, -1.0 / 6.0);
  ps = fma(d2, ps, 1.0);
  const double sd = d * ps;
  double pc = fma(d2, -1.0 / 720.0, 1.0 / 24.0);
  pc = fma(d2, pc, -0.5);
  const double cd = fma(d2, pc, 1.0);
  const double ca = SC_TAB[k & 255u][0], sa = SC_TAB[k & 255u][1];
  *c = fma(ca, cd, -(sa * sd));
  *s = fma(sa, cd, ca * sd);
}

/* -2 ln u, u = (n40 + 1) * 2^-40 in (0, 1]; never negative */
double orc_neg2ln40(uint64_t n40) {
  if (!noise_tabs_ready) orc_noise_tables(NULL, NULL);
  const double u = (double)(n40 + 1) * 9.094947017729282e-13; /* 2^-40, exact */
  uint64_t bits;
  memcpy(&bits, &u, 8);
  const int e = (int)((bits >> 52) & 0x7ff) - 1023;
  const int j = (int)((bits >> 45) & 127);
  const uint64_t mb = (bits & 0x000fffffffffffffull) | 0x3ff0000000000000ull;
  double m;
  memcpy(&m, &mb, 8);
  const double r = fma(m, LN_TAB[j][0], -1.0);
  double p = fma(r, -1.0 / 6.0, 0.2);
  p = fma(r, p, -0.25);
  p = fma(r, p, 1.0 / 3.0);
  p = fma(r, p, -0.5);
  p = fma(r, p, 1.0);
  p = r * p;
  const double ln = fma((double)e, 0.6931471805599453, LN_TAB[j][1]) + p;
  return fmax(-2.0 * ln, 0.0);
}

/* N(0,1) from two words: u1 from a and the top byte of b (40 bits), the angle from b's other 24 bits */
double orc_normal40(uint32_t a, uint32_t b) {
  const uint64_t n40 = (uint64_t)a | ((uint64_t)(b >> 24) << 32);
  double c, s;
  orc_unit2(b << 8, &c, &s);
  return sqrt(orc_neg2ln40(n40)) * c;
}

/* uniform point on the unit sphere from two words */
void orc_sphere(uint32_t a, uint32_t b, double *o) {
  double c, s;
  orc_unit2(a, &c, &s);
  const double z = 1.0 - ((double)b + 0.5) * 4.656612873077393e-10; /* 2^-31 */
  const double q = sqrt(fmax(fma(-z, z, 1.0), 0.0));
  o[0] = q * c;
  o[1] = q * s;
  o[2] = z;
}

/* test helper: the three draws of n blocks of four words */
void orc_noise_draws(const uint32_t *w, uint64_t n, double *circle, double *sph, double *nrm) {
  for (uint64_t i = 0; i < n; ++i) {
    orc_unit2(w[4 * i], &circle[2 * i], &circle[2 * i + 1]);
    orc_sphere(w[4 * i], w[4 * i + 1], sph + 3 * i);
    nrm[i] = orc_normal40(w[4 * i + 2], w[4 * i + 3]);
  }
}

static void noise_block(uint64_t seed, uint32_t stream, uint64_t index, uint32_t slot, uint32_t *o) {
  uint32_t ctr[4] = {(uint32_t)index, (uint32_t)(index >> 32), stream, slot};
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  orc_philox4x32_10(ctr, key, o);
}

enum { ST_DRIFT_CAM = 1, ST_DRIFT_PT = 2, ST_NOISE_CAM = 3, ST_NOISE_PT = 4, ST_NOISE_OBS = 5 };

/* src/baproblem.rs:282-289 (sequential fold a + b/num: cameras' centres, then points) */
void orc_mean(const double *cams, uint64_t C, const double *pts, uint64_t P, double *m) {
  double num = (double)(C + P);
  m[0] = m[1] = m[2] = 0.0;
  for (uint64_t i = 0; i < C; ++i) {
    double c[3];
    orc_center(cams + ORC_CAM_STRIDE * i, c);
    for (int k = 0; k < 3; ++k) m[k] = m[k] + c[k] / num;
  }
  for (uint64_t i = 0; i < P; ++i)
    for (int k = 0; k < 3; ++k) m[k] = m[k] + pts[3 * i + k] / num;
}

/* src/baproblem.rs:292-304 */
void orc_std(const double *cams, uint64_t C, const double *pts, uint64_t P, double *s) {
  double num = (double)(C + P), mean[3], acc[3] = {0, 0, 0};
  orc_mean(cams, C, pts, P, mean);
  for (uint64_t i = 0; i < C; ++i) {
    double c[3];
    orc_center(cams + ORC_CAM_STRIDE * i, c);
    for (int k = 0; k < 3; ++k) acc[k] = acc[k] + (c[k] - mean[k]) * (c[k] - mean[k]);
  }
  for (uint64_t i = 0; i < P; ++i)
    for (int k = 0; k < 3; ++k)
      acc[k] = acc[k] + (pts[3 * i + k] - mean[k]) * (pts[3 * i + k] - mean[k]);
  for (int k = 0; k < 3; ++k) s[k] = sqrt(acc[k] / num);
}

/* src/noise.rs:68-116 */
void orc_add_drift(double *cams, uint64_t C, double *pts, uint64_t P, double strength,
                   double angle_strength, double std, const double *dir, uint64_t seed) {
  /* origin = element nearest the world origin; fold1 keeps x only if strictly nearer (:80-86) */
  double origin[3] = {0, 0, 0}, best = 0.0;
  int have = 0;
  for (uint64_t i = 0; i < C + P; ++i) {
    double e[3];
    if (i < C)
      orc_center(cams + ORC_CAM_STRIDE * i, e);
    else
      memcpy(e, pts + 3 * (i - C), 24);
    double dd = mag3(e);
    if (!have || !(best < dd)) {
      best = dd;
      memcpy(origin, e, 24);
      have = 1;
    }
  }
  for (uint64_t i = 0; i < C; ++i) {
    double *cam = cams + ORC_CAM_STRIDE * i, c[3], z[2], R[9], dl[3], out[ORC_CAM_STRIDE];
    orc_center(cam, c);
    double d[3] = {c[0] - origin[0], c[1] - origin[1], c[2] - origin[2]};
    double distance = mag3(d);
    uint32_t o[4];
    noise_block(seed, ST_DRIFT_CAM, i, 0, o);
    z[0] = orc_normal40(o[0], o[1]); /* the angle's draw first (:104-107) */
    z[1] = orc_normal40(o[2], o[3]);
    double v1 = 1.0 + std * z[0], v2 = 1.0 + std * z[1];
    double angle = (angle_strength * v1) * pow(distance, 1.2); /* :97, drawn first (:105) */
    for (int k = 0; k < 3; ++k) dl[k] = (((dir[k] * strength) * v2) * distance) * distance; /* :92 */
    orc_from_angle_x(angle, R);
    orc_transform(cam, R, dl, out);
    memcpy(cam, out, sizeof out);
  }
  for (uint64_t i = 0; i < P; ++i) {
    double *p = pts + 3 * i, z[2];
    double d[3] = {p[0] - origin[0], p[1] - origin[1], p[2] - origin[2]};
    double distance = mag3(d);
    uint32_t o[4];
    noise_block(seed, ST_DRIFT_PT, i, 0, o);
    z[0] = orc_normal40(o[0], o[1]);
    double v = 1.0 + std * z[0];
    for (int k = 0; k < 3; ++k) p[k] = p[k] + (((dir[k] * strength) * v) * distance) * distance;
  }
}

/* src/noise.rs:47-56 */
void orc_add_drift_normalized(double *cams, uint64_t C, double *pts, uint64_t P, double strength,
                              double angle_strength, double std, uint64_t seed) {
  double s[3], dir[3];
  orc_std(cams, C, pts, P, s);
  normalize3(s, dir);
  double bal_std = mag3(s);
  orc_add_drift(cams, C, pts, P, strength * bal_std, angle_strength, std, dir, seed);
}

/* src/noise.rs:119-177.  Draw order of the reference per camera: axis, angle, translation direction,
 * magnitude (:140-141) = block 0 {sphere, N}, block 1 {sphere, N}; per point: direction, magnitude (:149) =
 * one block {sphere, N}; per observation: direction (nx, ny), r (:159-163) = one block {circle, -, N}. */
void orc_add_noise(double *cams, uint64_t C, double *pts, uint64_t P, double *uv, uint64_t O,
                   double translation_std, double rotation_std, double point_std,
                   double observations_std, uint64_t seed) {
  double s[3];
  orc_std(cams, C, pts, P, s);
  double bal_std = mag3(s);
  for (uint64_t i = 0; i < C; ++i) {
    double *cam = cams + ORC_CAM_STRIDE * i, ax[3], tr[3], R[9], dl[3];
    double out[ORC_CAM_STRIDE];
    uint32_t o0[4], o1[4];
    noise_block(seed, ST_NOISE_CAM, i, 0, o0);
    noise_block(seed, ST_NOISE_CAM, i, 1, o1);
    orc_sphere(o0[0], o0[1], ax);
    double angle = 0.0 + rotation_std * orc_normal40(o0[2], o0[3]);
    orc_sphere(o1[0], o1[1], tr);
    double mag = 0.0 + translation_std * orc_normal40(o1[2], o1[3]);
    orc_from_axis_angle(ax, angle, R);
    for (int k = 0; k < 3; ++k) dl[k] = (tr[k] * bal_std) * mag;
    orc_transform(cam, R, dl, out);
    memcpy(cam, out, sizeof out);
  }
  for (uint64_t i = 0; i < P; ++i) {
    double *p = pts + 3 * i, ax[3];
    uint32_t o[4];
    noise_block(seed, ST_NOISE_PT, i, 0, o);
    orc_sphere(o[0], o[1], ax);
    double mag = 0.0 + point_std * orc_normal40(o[2], o[3]);
    for (int k = 0; k < 3; ++k) p[k] = p[k] + ax[k] * mag;
  }
  for (uint64_t i = 0; i < O; ++i) {
    uint32_t o[4];
    double nx, ny;
    noise_block(seed, ST_NOISE_OBS, i, 0, o);
    orc_unit2(o[0], &nx, &ny);
    double r = 0.0 + observations_std * orc_normal40(o[2], o[3]);
    uv[2 * i] = uv[2 * i] + nx * r;
    uv[2 * i + 1] = uv[2 * i + 1] + ny * r;
  }
}

/* src/generate.rs:356-420 with the candidate stream of the CUDA path: candidate k draws the
 * triangle variate from Philox(seed, stream 16, index k, slot 0) and (rx, ry) from slot 1; accepted
 * candidates are kept in order.  Returns the number of points written (num_points), or 0 when more
 * than 10 * num_points candidates were rejected first (the reference panics). */
uint64_t orc_generate_world_points_uniform(const float *xyz, uint64_t nv, const uint32_t *tri, uint64_t nt,
                                           const double *cams, uint64_t C, uint64_t num_points,
                                           double max_dist, uint64_t seed, double *out) {
  if (!C || !nt || !num_points) return 0;
  double *cdf = (double *)malloc(8 * nt), *cen = (double *)malloc(24 * C);
  double run = 0.0;
  for (uint64_t i = 0; i < nt; ++i) {
    double v[3][3], e1[3], e2[3], n[3];
    for (int j = 0; j < 3; ++j)
      for (int k = 0; k < 3; ++k) v[j][k] = (double)xyz[3 * (uint64_t)tri[3 * i + j] + k];
    for (int k = 0; k < 3; ++k) {
      e1[k] = v[1][k] - v[0][k];
      e2[k] = v[2][k] - v[0][k];
    }
    cross3(e1, e2, n);
    run += mag3(n) / 2.0;
    cdf[i] = run;
  }
  for (uint64_t i = 0; i < C; ++i) orc_center(cams + ORC_CAM_STRIDE * i, cen + 3 * i);
  const double max_d2 = max_dist * max_dist;
  uint64_t got = 0, fails = 0;
  for (uint64_t k = 0; got < num_points && fails < 10 * num_points && run > 0.0; ++k) {
    uint32_t c0[4] = {(uint32_t)k, (uint32_t)(k >> 32), 16u, 0u}, c1[4] = {(uint32_t)k, (uint32_t)(k >> 32), 16u, 1u};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)}, o[4], q[4];
    orc_philox4x32_10(c0, key, o);
    orc_philox4x32_10(c1, key, q);
    double u = (double)((((uint64_t)o[0]) | ((uint64_t)o[1] << 32)) >> 11) * 1.1102230246251565e-16;
    double rx = (double)((((uint64_t)q[0]) | ((uint64_t)q[1] << 32)) >> 11) * 1.1102230246251565e-16;
    double ry = (double)((((uint64_t)q[2]) | ((uint64_t)q[3] << 32)) >> 11) * 1.1102230246251565e-16;
    double target = u * run;
    uint64_t lo = 0, hi = nt - 1;
    while (lo < hi) {
      uint64_t mid = (lo + hi) >> 1;
      if (cdf[mid] > target) hi = mid; else lo = mid + 1;
    }
    if (rx + ry > 1.0) {
      rx = 1.0 - rx;
      ry = 1.0 - ry;
    }
    double p[3];
    for (int d = 0; d < 3; ++d) {
      double v0 = (double)xyz[3 * (uint64_t)tri[3 * lo] + d], v1 = (double)xyz[3 * (uint64_t)tri[3 * lo + 1] + d],
             v2 = (double)xyz[3 * (uint64_t)tri[3 * lo + 2] + d];
      p[d] = (v0 + rx * (v1 - v0)) + ry * (v2 - v0);
    }
    int near = 0;
    for (uint64_t c = 0; c < C && !near; ++c) {
      double dx = cen[3 * c] - p[0], dy = cen[3 * c + 1] - p[1], dz = cen[3 * c + 2] - p[2];
      if ((dx * dx + dy * dy) + dz * dz <= max_d2) near = 1;
    }
    if (near) {
      memcpy(out + 3 * got, p, 24);
      ++got;
    } else {
      ++fails;
    }
  }
  free(cdf);
  free(cen);
  return got == num_points ? got : 0;
}

/* src/noise.rs:388-416 with BAProblem::extent / dimensions (src/baproblem.rs:307-337) */
void orc_add_sin_noise(double *cams, uint64_t C, double *pts, uint64_t P, const double *dir,
                       const double *noise_dir, double strength, double frequency) {
  double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (uint64_t i = 0; i < C; ++i) {
    double c[3];
    orc_center(cams + ORC_CAM_STRIDE * i, c);
    for (int k = 0; k < 3; ++k) {
      lo[k] = fmin(lo[k], c[k]);
      hi[k] = fmax(hi[k], c[k]);
    }
  }
  for (uint64_t i = 0; i < P; ++i)
    for (int k = 0; k < 3; ++k) {
      lo[k] = fmin(lo[k], pts[3 * i + k]);
      hi[k] = fmax(hi[k], pts[3 * i + k]);
    }
  double dim[3], nd[3];
  for (int k = 0; k < 3; ++k) {
    dim[k] = hi[k] - lo[k];
    if (dim[k] == 0.0) dim[k] = 1e-8; /* "Add epsilon to nonexistent dimensions" */
  }
  normalize3(noise_dir, nd);
  static const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  for (uint64_t i = 0; i < C + P; ++i) {
    double x[3];
    if (i < C)
      orc_center(cams + ORC_CAM_STRIDE * i, x);
    else
      memcpy(x, pts + 3 * (i - C), 24);
    double q[3] = {x[0] / dim[0], x[1] / dim[1], x[2] / dim[2]};
    double dot = (q[0] * dir[0] + q[1] * dir[1]) + q[2] * dir[2];
    double amp = sin(dot * frequency * 3.14159265358979323846) * strength;
    double d[3] = {nd[0] * amp, nd[1] * amp, nd[2] * amp};
    if (i < C) {
      double out[ORC_CAM_STRIDE];
      orc_transform(cams + ORC_CAM_STRIDE * i, I, d, out);
      memcpy(cams + ORC_CAM_STRIDE * i, out, sizeof out);
    } else {
      for (int k = 0; k < 3; ++k) pts[3 * (i - C) + k] = x[k] + d[k];
    }
  }
}

/* src/baproblem.rs:265-279 */
double orc_total_reprojection_error(const double *cams, uint64_t C, const double *pts,
                                    const uint64_t *offsets, const uint64_t *point_idx,
                                    const double *uv, double norm) {
  double total = 0.0;
  for (uint64_t c = 0; c < C; ++c) {
    double s = 0.0;
    for (uint64_t i = offsets[c]; i < offsets[c + 1]; ++i) {
      double pc[3], q[2];
      orc_project_world(cams + ORC_CAM_STRIDE * c, pts + 3 * point_idx[i], pc);
      orc_project(cams + ORC_CAM_STRIDE * c, pc, q);
      s += pow(fabs(q[0] - uv[2 * i]), norm) + pow(fabs(q[1] - uv[2 * i + 1]), norm);
    }
    total += s;
  }
  return pow(total, 1. / norm);
}
