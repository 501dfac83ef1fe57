/* cpu_ref.c — multithreaded CPU arm of the ORACLE (test infrastructure; see c2b_oracle.h).
 *
 * This is the stand-in for the reference binary, which cannot be built in this image (no
 * rustc/cargo, no libembree3, no network).  It restates the reference ALGORITHM, not the GPU
 * design: an OpenMP `parallel for schedule(dynamic)` over cameras (rayon par_iter,
 * src/generate.rs:435), inside it the brute-force loop over ALL points with the reference's
 * predicate order (src/generate.rs:446-469), a per-camera ray batch, any-hit against a CPU
 * BVH (binned SAH BVH2 — the role Embree's BVH plays at src/generate.rs:472) using exactly the
 * oracle's f32 watertight predicate, and per-camera output vectors (src/generate.rs:473-478).
 * Because any-hit is an OR over triangles and the box test is conservative, the result is
 * identical to the brute-force oracle (checked in tests/test_oracle_visibility.py).
 *
 * Single translation unit: includes c2b_oracle.c.
 */
#include "c2b_oracle.c"

#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
  float lo[3], hi[3];
  uint32_t left;  /* internal: index of left child (right = left+1); leaf: first triangle */
  uint32_t count; /* 0 => internal */
} bnode;

typedef struct {
  bnode *nodes;
  uint32_t n_nodes;
  tri9 *tris; /* reordered */
  uint64_t n_tris;
} cpu_bvh;

static void tri_bounds(const tri9 *t, float *lo, float *hi) {
  for (int k = 0; k < 3; ++k) {
    lo[k] = fminf(t->v[k], fminf(t->v[3 + k], t->v[6 + k]));
    hi[k] = fmaxf(t->v[k], fmaxf(t->v[3 + k], t->v[6 + k]));
  }
}

static float half_area(const float *lo, const float *hi) {
  float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
  return dx * dy + dy * dz + dz * dx;
}

static void build_rec(cpu_bvh *b, uint32_t ni, uint64_t first, uint64_t count) {
  bnode *nd = &b->nodes[ni];
  float clo[3] = {INFINITY, INFINITY, INFINITY}, chi[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int k = 0; k < 3; ++k) {
    nd->lo[k] = INFINITY;
    nd->hi[k] = -INFINITY;
  }
  for (uint64_t i = first; i < first + count; ++i) {
    float lo[3], hi[3];
    tri_bounds(&b->tris[i], lo, hi);
    for (int k = 0; k < 3; ++k) {
      nd->lo[k] = fminf(nd->lo[k], lo[k]);
      nd->hi[k] = fmaxf(nd->hi[k], hi[k]);
      float c = 0.5f * (lo[k] + hi[k]);
      clo[k] = fminf(clo[k], c);
      chi[k] = fmaxf(chi[k], c);
    }
  }
  if (count <= 4) {
    nd->left = (uint32_t)first;
    nd->count = (uint32_t)count;
    return;
  }
  /* binned SAH, 16 bins on every axis */
  enum { NB = 16 };
  int best_axis = -1, best_split = 0;
  float best_cost = INFINITY;
  for (int ax = 0; ax < 3; ++ax) {
    float ext = chi[ax] - clo[ax];
    if (!(ext > 0.0f)) continue;
    float blo[NB][3], bhi[NB][3];
    uint64_t bc[NB];
    for (int j = 0; j < NB; ++j) {
      bc[j] = 0;
      for (int k = 0; k < 3; ++k) {
        blo[j][k] = INFINITY;
        bhi[j][k] = -INFINITY;
      }
    }
    for (uint64_t i = first; i < first + count; ++i) {
      float lo[3], hi[3];
      tri_bounds(&b->tris[i], lo, hi);
      int j = (int)((0.5f * (lo[ax] + hi[ax]) - clo[ax]) / ext * NB);
      if (j >= NB) j = NB - 1;
      if (j < 0) j = 0;
      bc[j]++;
      for (int k = 0; k < 3; ++k) {
        blo[j][k] = fminf(blo[j][k], lo[k]);
        bhi[j][k] = fmaxf(bhi[j][k], hi[k]);
      }
    }
    float rarea[NB];
    uint64_t rcnt[NB];
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    uint64_t cnt = 0;
    for (int j = NB - 1; j > 0; --j) {
      for (int k = 0; k < 3; ++k) {
        lo[k] = fminf(lo[k], blo[j][k]);
        hi[k] = fmaxf(hi[k], bhi[j][k]);
      }
      cnt += bc[j];
      rarea[j] = cnt ? half_area(lo, hi) : 0.0f;
      rcnt[j] = cnt;
    }
    for (int k = 0; k < 3; ++k) {
      lo[k] = INFINITY;
      hi[k] = -INFINITY;
    }
    cnt = 0;
    for (int j = 0; j < NB - 1; ++j) {
      for (int k = 0; k < 3; ++k) {
        lo[k] = fminf(lo[k], blo[j][k]);
        hi[k] = fmaxf(hi[k], bhi[j][k]);
      }
      cnt += bc[j];
      if (cnt == 0 || rcnt[j + 1] == 0) continue;
      float cost = half_area(lo, hi) * (float)cnt + rarea[j + 1] * (float)rcnt[j + 1];
      if (cost < best_cost) {
        best_cost = cost;
        best_axis = ax;
        best_split = j;
      }
    }
  }
  uint64_t mid;
  if (best_axis < 0) {
    mid = first + count / 2; /* all centroids coincide */
  } else {
    float ext = chi[best_axis] - clo[best_axis];
    uint64_t i = first, j = first + count;
    while (i < j) {
      float lo[3], hi[3];
      tri_bounds(&b->tris[i], lo, hi);
      int bin = (int)((0.5f * (lo[best_axis] + hi[best_axis]) - clo[best_axis]) / ext * NB);
      if (bin >= NB) bin = NB - 1;
      if (bin < 0) bin = 0;
      if (bin <= best_split) {
        ++i;
      } else {
        --j;
        tri9 tmp = b->tris[i];
        b->tris[i] = b->tris[j];
        b->tris[j] = tmp;
      }
    }
    mid = i;
    if (mid == first || mid == first + count) mid = first + count / 2;
  }
  uint32_t l = b->n_nodes;
  b->n_nodes += 2;
  nd->left = l;
  nd->count = 0;
  build_rec(b, l, first, mid - first);
  build_rec(b, l + 1, mid, first + count - mid);
}

static cpu_bvh *bvh_build(tri9 *tris, uint64_t n) {
  cpu_bvh *b = (cpu_bvh *)calloc(1, sizeof(cpu_bvh));
  b->tris = tris;
  b->n_tris = n;
  b->nodes = (bnode *)malloc(sizeof(bnode) * (2 * (n ? n : 1)));
  b->n_nodes = 1;
  if (n)
    build_rec(b, 0, 0, n);
  else
    b->n_nodes = 0;
  return b;
}

/* conservative slab test: boxes are inflated relative to their distance from the origin and the
 * interval compare carries slack, so the watertight triangle test (whose rounding error is a few
 * f32 ulps of the same magnitudes) never accepts a hit outside a rejected box */
static int box_hit(const orc_ray *r, const float *inv, const bnode *nd) {
  float tmin = 0.0f, tmax = r->tfar;
  for (int k = 0; k < 3; ++k) {
    float a = nd->lo[k] - r->org[k], b = nd->hi[k] - r->org[k];
    float e = 4e-6f * fmaxf(fabsf(a), fabsf(b)) + 1e-30f;
    a -= e;
    b += e;
    float t0 = a * inv[k], t1 = b * inv[k];
    tmin = fmaxf(tmin, fminf(t0, t1));
    tmax = fminf(tmax, fmaxf(t0, t1));
  }
  return tmin <= tmax * 1.00001f + 1e-30f;
}

static int bvh_occluded(const cpu_bvh *b, const orc_ray *r) {
  if (!b->n_nodes) return 0;
  if (!(r->tfar >= 0.0f)) return 0; /* NaN or negative tfar: nothing can satisfy 0 < t <= tfar */
  float inv[3] = {1.0f / r->dir[0], 1.0f / r->dir[1], 1.0f / r->dir[2]};
  uint32_t stack[128];
  int sp = 0;
  stack[sp++] = 0;
  while (sp) {
    const bnode *nd = &b->nodes[stack[--sp]];
    if (!box_hit(r, inv, nd)) continue;
    if (nd->count) {
      for (uint32_t i = 0; i < nd->count; ++i) {
        const tri9 *t = &b->tris[nd->left + i];
        if (tri_test(r, t->v, t->v + 3, t->v + 6, 0, 0, 0, 0)) return 1;
      }
    } else if (sp + 2 <= 128) {
      stack[sp++] = nd->left;
      stack[sp++] = nd->left + 1;
    }
  }
  return 0;
}

orc_vis *orc_ref_visibility_graph(const float *xyz, uint64_t nv, const uint32_t *tri, uint64_t nt,
                                  const double *cams, uint64_t C, const double *pts, uint64_t P,
                                  double max_dist, int endpoint_guard_rel, int n_threads,
                                  int *threads_used) {
  uint64_t ntri = 0;
  tri9 *T = gather_tris(xyz, nv, tri, nt, &ntri);
  cpu_bvh *bvh = bvh_build(T, ntri);
  candvec *per_cam = (candvec *)calloc(C ? C : 1, sizeof(candvec));
  int used = 1;
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#pragma omp parallel
  {
#pragma omp single
    used = omp_get_num_threads();
  }
#endif
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t c = 0; c < (int64_t)C; ++c) {
    const double *cam = cams + ORC_CAM_STRIDE * c;
    double center[3];
    orc_center(cam, center);
    /* src/generate.rs:444-469: obs + ray vectors filled by the loop over all points */
    candvec obs = {0};
    orc_ray *rays = 0;
    uint64_t nr = 0, capr = 0;
    for (uint64_t i = 0; i < P; ++i) {
      double uv[2];
      if (!cull_pair(cam, center, pts + 3 * i, max_dist, uv, 0)) continue;
      if (nr == capr) {
        capr = capr ? capr * 2 : 256;
        rays = (orc_ray *)realloc(rays, capr * sizeof(orc_ray));
      }
      orc_make_ray(center, pts + 3 * i, &rays[nr]);
      if (endpoint_guard_rel) rays[nr].tfar = rays[nr].tfar * (1.0f - 3.814697265625e-06f);
      ++nr;
      cv_push(&obs, i, uv, 0, 0);
    }
    /* src/generate.rs:472-478: occlusion stream, keep rays that were not hit */
    uint64_t w = 0;
    for (uint64_t j = 0; j < nr; ++j) {
      if (bvh_occluded(bvh, &rays[j])) continue;
      obs.pt[w] = obs.pt[j];
      obs.uv[2 * w] = obs.uv[2 * j];
      obs.uv[2 * w + 1] = obs.uv[2 * j + 1];
      ++w;
    }
    obs.n = w;
    free(rays);
    per_cam[c] = obs;
  }
  if (threads_used) *threads_used = used;
  orc_vis *r = (orc_vis *)calloc(1, sizeof(orc_vis));
  r->n_cameras = C;
  r->offsets = (uint64_t *)calloc(C + 1, 8);
  uint64_t total = 0;
  for (uint64_t c = 0; c < C; ++c) {
    r->offsets[c] = total;
    total += per_cam[c].n;
  }
  r->offsets[C] = total;
  r->n_obs = total;
  r->point_idx = (uint64_t *)malloc(8 * (total ? total : 1));
  r->uv = (double *)malloc(16 * (total ? total : 1));
  for (uint64_t c = 0; c < C; ++c) {
    memcpy(r->point_idx + r->offsets[c], per_cam[c].pt, 8 * per_cam[c].n);
    memcpy(r->uv + 2 * r->offsets[c], per_cam[c].uv, 16 * per_cam[c].n);
    free(per_cam[c].pt);
    free(per_cam[c].uv);
    free(per_cam[c].occ);
    free(per_cam[c].flg);
  }
  free(per_cam);
  free(bvh->nodes);
  free(bvh);
  free(T);
  return r;
}
