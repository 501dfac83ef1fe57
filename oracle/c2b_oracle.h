/* c2b_oracle.h — CPU ORACLE for the city2ba visibility / observation / noise hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under city2ba_b200/ (the product) may include,
 * link, import or execute this code; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs do, and only as the checker or as
 * the timed CPU arm.
 *
 * It is a plain-C restatement of the reference's algorithm (tkonolige/city2ba, Rust).
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference).  Pinning status:
 *   - camera math: pinned against the reference's four known-answer tests
 *     (src/baproblem.rs:64-75, 227-249) in tests/test_oracle_camera.py.
 *   - occlusion: the reference calls Intel Embree 3.8.0 through embree-rs 0.3.6
 *     (Cargo.lock:278-279); neither Rust nor Embree exists in this image and the
 *     reference holds no golden visibility vectors, so this boundary is
 *     PARITY UNPINNED.  The oracle fixes ONE fully specified f32 predicate
 *     (watertight edge-function test in the form of Embree's robust "Pluecker"
 *     intersector on origin-relative vertices, fixed operation order, explicit
 *     fmaf only, no contraction) and additionally reports the rays that lie within a stated
 *     epsilon of an edge / a grazing face / the ray end point, where Embree's
 *     rounding could legitimately differ.
 *   - noise: the reference draws from rand 0.6.5 thread_rng() (unseedable); the
 *     oracle and the CUDA path share a Philox4x32-10 stream instead, so parity is
 *     exact between them and distributional against the reference.
 */
#ifndef C2B_ORACLE_H
#define C2B_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* camera record: 15 doubles = R column-major [9] (cgmath Matrix3 {x,y,z} columns),
 * t [3], intrinsics f,k1,k2 [3]  (src/baproblem.rs:131-138) */
#define ORC_CAM_STRIDE 15

/* ---- camera math (src/baproblem.rs:78-225) ---- */
void orc_project_world(const double *cam, const double *p, double *out3);
void orc_project(const double *cam, const double *pc, double *out2);
void orc_center(const double *cam, double *out3);
void orc_to_world(const double *cam, const double *p, double *out3);
void orc_from_position_direction(const double *pos3, const double *R9, double *cam_out);
void orc_transform(const double *cam, const double *dR9, const double *dloc3, double *cam_out);
void orc_from_rodrigues(const double *v3, double *R9);
void orc_to_rodrigues(const double *R9, double *v3);
void orc_from_vec(const double *v9, double *cam_out);
void orc_to_vec(const double *cam, double *v9);
void orc_from_angle_x(double rad, double *R9);
void orc_from_angle_y(double rad, double *R9);
void orc_from_axis_angle(const double *axis3, double rad, double *R9);
double orc_deg_to_rad(double deg);

/* ---- ray / triangle predicate (replaces Embree rtcOccluded1M, src/generate.rs:472) ---- */
typedef struct {
  float org[3];
  float dir[3];
  float tfar;
} orc_ray;
void orc_make_ray(const double *center3, const double *point3, orc_ray *ray);
/* returns 1 if the triangle (three f32 xyz vertices) occludes the ray: 0 < t <= tfar */
int orc_ray_triangle(const orc_ray *ray, const float *v0, const float *v1, const float *v2);

/* flag bits for rays within epsilon of a decision boundary */
#define ORC_FLAG_EDGE 1u
#define ORC_FLAG_GRAZE 2u
#define ORC_FLAG_ENDPOINT 4u
#define ORC_FLAG_CULL 8u /* cull-stage compare within 1e-12 relative of its threshold */

/* ---- visibility graph (src/generate.rs:424-481), brute force over triangles ---- */
typedef struct {
  uint64_t n_cameras;
  uint64_t n_candidates;  /* pairs that passed the cull + frustum (rays cast) */
  uint64_t n_obs;         /* candidates that were not occluded */
  uint64_t *cand_offsets; /* [C+1] */
  uint64_t *cand_point;   /* [n_candidates] ascending point index per camera */
  double *cand_uv;        /* [2*n_candidates] */
  uint8_t *cand_occluded; /* [n_candidates] */
  uint8_t *cand_flags;    /* [n_candidates] ORC_FLAG_* (0 when flags were not requested) */
  uint64_t *offsets;      /* [C+1] CSR of visible observations */
  uint64_t *point_idx;    /* [n_obs] */
  double *uv;             /* [2*n_obs] */
  uint64_t n_flag_edge, n_flag_graze, n_flag_endpoint, n_flag_cull;
} orc_vis;

/* xyz: 3*nv f32; tri: 3*nt u32 (triples with a repeated index are ignored, as is any
 * triple that touches an out-of-range vertex).  endpoint_guard_rel: 0 = reference
 * behaviour tfar = f32(|d|) - 1e-6f; 1 = tfar additionally scaled by (1 - 2^-18).
 * want_flags: compute the flagged-epsilon classes (slower). */
orc_vis *orc_visibility_graph(const float *xyz, uint64_t nv, const uint32_t *tri, uint64_t nt,
                              const double *cams, uint64_t C, const double *pts, uint64_t P,
                              double max_dist, int endpoint_guard_rel, int want_flags);
void orc_vis_free(orc_vis *v);

/* A/B counter: Embree 3's default Moeller-Trumbore triangle test, restated (unpinned), and the occluded
 * verdict of every candidate (cand_offsets / cand_point of an orc_vis) under it, one byte each */
int orc_ray_triangle_mt(const orc_ray *ray, const float *v0, const float *v1, const float *v2);
void orc_occluded_mt(const float *xyz, uint64_t nv, const uint32_t *tri, uint64_t nt, const double *cams,
                     uint64_t C, const double *pts, const uint64_t *cand_offsets, const uint64_t *cand_point,
                     int endpoint_guard_rel, uint8_t *occluded);

/* ---- multithreaded CPU reference arm: same predicates, BVH-accelerated any-hit,
 *      OpenMP over cameras (the rayon par_iter of src/generate.rs:435).  Returns
 *      visible CSR only (cand_* are NULL).  n_threads<=0: all cores. ---- */
orc_vis *orc_ref_visibility_graph(const float *xyz, uint64_t nv, const uint32_t *tri, uint64_t nt,
                                  const double *cams, uint64_t C, const double *pts, uint64_t P,
                                  double max_dist, int endpoint_guard_rel, int n_threads,
                                  int *threads_used);

/* ---- synthetic lattice (src/synthetic.rs:163-258, 313-344) ---- */
uint64_t orc_grid_num_cameras(uint64_t cpb, uint64_t n_blocks);
uint64_t orc_grid_num_points(uint64_t ppb, uint64_t n_blocks);
void orc_grid_cameras(uint64_t cpb, uint64_t n_blocks, double block_length, double camera_height,
                      double *cams_out);
void orc_grid_points(uint64_t ppb, uint64_t n_blocks, double block_length, double block_inset,
                     double point_height, double *pts_out);
void orc_line_cameras(uint64_t num_cameras, double length, double camera_height, double *cams_out);
void orc_line_points(uint64_t num_points, double length, double point_offset, double point_height,
                     double *pts_out);
/* analytic 2-D wall occlusion of `synthetic` (src/synthetic.rs:52-124) incl. the :93 quirk */
int orc_hits_building(const double *c3, const double *p3, double block_length, double block_inset);
/* synthetic-mode visibility (src/synthetic.rs:268-297): analytic!=0 uses hits_building,
 * analytic==0 skips occlusion (synthetic_line :353-379).  Observations are emitted in ascending
 * point index (the reference's rstar order is implementation defined; compare as sets). */
orc_vis *orc_synthetic_visibility(const double *cams, uint64_t C, const double *pts, uint64_t P,
                                  double max_dist, int analytic, double block_length,
                                  double block_inset);

/* ---- city-block box mesh (new artefact): one box per block, 8 vertices,
 *      12 triangles; footprint [b*L+inset,(b+1)*L-inset]^2, y in [0,H] ---- */
void orc_city_mesh(uint64_t n_blocks, double block_length, double block_inset, double height,
                   float *xyz_out /*3*8*n^2*/, uint32_t *tri_out /*3*12*n^2*/);

/* ---- noise (src/noise.rs:35-177; src/baproblem.rs:282-304) with a Philox4x32-10 stream ---- */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* the draws (c2b_oracle.c "the noise draws"): tables, (cos, sin)(2 pi w / 2^32), -2 ln((n40 + 1) 2^-40),
 * N(0,1) from two words, a uniform point on the sphere from two words */
void orc_noise_tables(double *sc /*512 or NULL*/, double *ln /*256 or NULL*/);
void orc_unit2(uint32_t w, double *c, double *s);
double orc_neg2ln40(uint64_t n40);
double orc_normal40(uint32_t a, uint32_t b);
void orc_sphere(uint32_t a, uint32_t b, double *out3);
void orc_noise_draws(const uint32_t *words, uint64_t n, double *circle2, double *sphere3, double *normal1);
void orc_mean(const double *cams, uint64_t C, const double *pts, uint64_t P, double *out3);
void orc_std(const double *cams, uint64_t C, const double *pts, uint64_t P, double *out3);
void orc_add_drift(double *cams, uint64_t C, double *pts, uint64_t P, double strength,
                   double angle_strength, double std, const double *dir3, uint64_t seed);
void orc_add_drift_normalized(double *cams, uint64_t C, double *pts, uint64_t P, double strength,
                              double angle_strength, double std, uint64_t seed);
void orc_add_noise(double *cams, uint64_t C, double *pts, uint64_t P, double *uv, uint64_t O,
                   double translation_std, double rotation_std, double point_std,
                   double observations_std, uint64_t seed);
/* total_reprojection_error (src/baproblem.rs:265-279) on CSR observations */
uint64_t orc_generate_world_points_uniform(const float *xyz, uint64_t nv, const uint32_t *tri, uint64_t nt,
                                           const double *cams, uint64_t C, uint64_t num_points,
                                           double max_dist, uint64_t seed, double *out);
void orc_add_sin_noise(double *cams, uint64_t C, double *pts, uint64_t P, const double *dir,
                       const double *noise_dir, double strength, double frequency);
double orc_total_reprojection_error(const double *cams, uint64_t C, const double *pts,
                                    const uint64_t *offsets, const uint64_t *point_idx,
                                    const double *uv, double norm);

#ifdef __cplusplus
}
#endif
#endif
