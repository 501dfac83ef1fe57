// tests/cpp/test_host.cpp — the reference's own tests restated against the C++ host mirror
// (include/city2ba.hpp).  Mirrors: the inline unit tests of src/baproblem.rs:64-75,227-249 and the
// library property tests of tests/main.rs:130-201 (every noise function must increase the total
// reprojection error of synthetic_grid(10,20,3,5,1,1,1,10); synthetic_line keeps > 20 cameras),
// plus BAL text / binary round trips (src/baproblem.rs:580-785).
//   test_host cpu            camera math, BAL I/O and cull on hand-made data (no GPU)
//   test_host gpu <out.bbal> everything; writes the culled synthetic grid for a cross-check in pytest
//   test_host obj <file.obj> dumps what tobj::load_obj read (cross-check against the Python loader)
#include <cstdio>
#include <cstdlib>
#include <set>
#include <string>

#include "city2ba.hpp"

using namespace city2ba;

static int failures = 0;
#define CHECK(cond)                                                        \
  do {                                                                     \
    if (!(cond)) {                                                         \
      std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);        \
      std::fflush(stdout);                                                 \
      ++failures;                                                          \
    }                                                                      \
  } while (0)

static double dist(const Vector3 &a, const Vector3 &b) {
  return std::sqrt((a[0] - b[0]) * (a[0] - b[0]) + (a[1] - b[1]) * (a[1] - b[1]) + (a[2] - b[2]) * (a[2] - b[2]));
}

// src/baproblem.rs:64-75
static void rodrigues_idempotent() {
  for (const Vector3 &v : {Vector3{1., 2., 3.}, Vector3{0., 0., 0.}, Vector3{-1.2, 0., 1.7}})
    CHECK(dist(to_rodrigues(from_rodrigues(v)), v) < 1e-10);
}
// src/baproblem.rs:227-234
static void test_project_world() {
  const Point3 p{0.0, 0.0, -1.0};
  const auto c = SnavelyCamera::from_vec({0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0});
  const Point3 pc = c.project_world(p);
  CHECK(pc[2] < 0.0);
  CHECK(pc[0] == 0.0 && pc[1] == 0.0);
}
// src/baproblem.rs:236-242
static void test_project() {
  const Point3 p{0.0, 0.0, -1.0};
  const auto c = SnavelyCamera::from_vec({0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0});
  const auto uv = c.project(c.project_world(p));
  CHECK(uv[0] == 0.0 && uv[1] == 0.0);
}
// src/baproblem.rs:244-249
static void test_project_isomorphic() {
  const Point3 p{1.0, 3.0, -1.0};
  const auto c = SnavelyCamera::from_vec({3.0, 5.0, -2.0, 0.5, -0.2, 0.1, 1.0, 0.0, 0.0});
  CHECK(dist(c.to_world(c.project_world(p)), p) < 1e-8);
}

static bool same_problem(const BAProblem &a, const BAProblem &b, double tol) {
  if (a.num_cameras() != b.num_cameras() || a.num_points() != b.num_points() ||
      a.num_observations() != b.num_observations())
    return false;
  for (size_t i = 0; i < a.num_points(); ++i)
    if (dist(a.points[i], b.points[i]) > tol) return false;
  for (size_t c = 0; c < a.num_cameras(); ++c) {
    const auto va = a.cameras[c].to_vec(), vb = b.cameras[c].to_vec();
    for (int k = 0; k < 9; ++k)
      if (std::fabs(va[k] - vb[k]) > 1e-9) return false;
    if (a.vis_graph[c].size() != b.vis_graph[c].size()) return false;
    for (size_t i = 0; i < a.vis_graph[c].size(); ++i)
      if (a.vis_graph[c][i].first != b.vis_graph[c][i].first ||
          a.vis_graph[c][i].second != b.vis_graph[c][i].second)
        return false;
  }
  return true;
}

// hand-made problem: an isolated camera, an unseen point, both file formats
static void graph_and_io(const std::string &tmp) {
  std::vector<SnavelyCamera> cams;
  for (int i = 0; i < 3; ++i)
    cams.push_back(SnavelyCamera::from_position_direction({double(i), 1.0, 0.25 * i}, from_angle_y(0.3 * i)));
  std::vector<Point3> pts;
  for (int i = 0; i < 6; ++i) pts.push_back({0.5 * i, 0.1 * i, -3.0 - 0.2 * i});
  VisGraph g(3);
  // cameras 0 and 1 both see points 0..4; camera 2 sees nothing; point 5 is seen by nobody
  for (int cam = 0; cam < 2; ++cam)
    for (size_t p = 0; p < 5; ++p) {
      const auto uv = cams[cam].project(cams[cam].project_world(pts[p]));
      g[cam].push_back({p, {uv[0], uv[1]}});
    }
  const BAProblem ba = BAProblem::from_visibility(cams, pts, g);
  CHECK(ba.num_observations() == 10);
  CHECK(ba.total_reprojection_error(2.0) < 1e-12);
  // largest_connected_component as the reference wrote it (src/baproblem.rs:456-534): the component
  // is {c0, c1, p0..p4}, but the observation filter at :523 looks up sets[point index] in the combined
  // (cameras, then points) numbering, so the observations of point 2 are tested against entity 2 =
  // camera 2, which is outside the component, and are dropped.
  const BAProblem lcc = ba.largest_connected_component();
  CHECK(lcc.num_cameras() == 2 && lcc.num_points() == 5 && lcc.num_observations() == 8);
  const BAProblem culled = ba.cull();  // point 2 is now unseen and goes; everything else stays
  CHECK(culled.num_cameras() == 2 && culled.num_points() == 4 && culled.num_observations() == 8);
  CHECK(culled.total_reprojection_error(1.0) < 1e-12);
  CHECK(culled.to_string() == "Bundle Adjustment Problem with 2 cameras, 4 points, and 8 observations");
  bool threw = false;
  try {
    BAProblem::from_visibility(cams, pts, VisGraph(2));
  } catch (const std::logic_error &) {
    threw = true;
  }
  CHECK(threw);  // assert!(cams.len() == obs.len())
  for (const char *ext : {".bal", ".bbal"}) {
    const std::string path = tmp + "/roundtrip" + ext;
    culled.write(path);
    CHECK(same_problem(BAProblem::from_file(path), culled, 1e-12));
  }
  CHECK(BAProblem::fmt(0.30000000000000004) == "0.30000000000000004" && BAProblem::fmt(1e-7) == "0.0000001" &&
        BAProblem::fmt(1.0) == "1" && BAProblem::fmt(-2.5e10) == "-25000000000");
  threw = false;
  try {
    culled.write(tmp + "/x.txt");
  } catch (const Error &e) {
    threw = e.kind == Error::IOError;
  }
  CHECK(threw);
}


// a small scene in the shape of the reference's tests/box.obj: a quad ground plane, a cube of quads,
// one pentagon (fan triangulation) and a poly-line object; written to disk and read back by tobj::load_obj
static std::string write_obj(const std::string &tmp) {
  const std::string path = tmp + "/scene.obj";
  std::ofstream f(path);
  f << "# test scene\nmtllib none.mtl\no Plane\n";
  f << "v -4 0 -4\nv 4 0 -4\nv 4 0 4\nv -4 0 4\nvn 0 1 0\nusemtl None\ns off\nf 1//1 2//1 3//1 4//1\n";
  f << "o Cube\n";
  for (int k = 0; k < 8; ++k) f << "v " << ((k & 1) ? 1 : -1) << ' ' << ((k & 2) ? 2 : 0) << ' ' << ((k & 4) ? 1 : -1) << "\n";
  // vertices 5..12; faces as quads, one of them with negative (relative) indices
  f << "f 5 6 8 7\nf 9 11 12 10\nf 5 9 10 6\nf 7 8 12 11\nf 5 7 11 9\nf -7 -5 -1 -3\n";
  f << "o Roof\nv -1 2 -1\nv 1 2 -1\nv 1.5 2 0\nv 1 2 1\nv -1 2 1\nf 13 14 15 16 17\n";
  f << "o Curve\n";
  for (int k = 0; k < 6; ++k) f << "v " << -3.0 + k << " 0.5 3\n";
  for (int k = 0; k < 5; ++k) f << "l " << 18 + k << ' ' << 19 + k << "\n";
  return path;
}

static void obj_and_cameras(const std::string &tmp) {
  const auto models = tobj::load_obj(write_obj(tmp));
  CHECK(models.size() == 4);
  if (models.size() != 4) return;
  CHECK(models[0].name == "Plane" && models[1].name == "Cube" && models[2].name == "Roof" && models[3].name == "Curve");
  CHECK(models[0].mesh.positions.size() == 12 && models[0].mesh.indices == (std::vector<uint32_t>{0, 1, 2, 0, 2, 3}));
  CHECK(models[1].mesh.positions.size() == 24 && models[1].mesh.indices.size() == 36);
  // pentagon: fan from its first vertex
  CHECK(models[2].mesh.indices == (std::vector<uint32_t>{0, 1, 2, 0, 2, 3, 0, 3, 4}));
  // line records: index pairs, vertices re-used between consecutive segments
  CHECK(models[3].mesh.positions.size() == 18 && models[3].mesh.indices == (std::vector<uint32_t>{0, 1, 1, 2, 2, 3, 3, 4, 4, 5}));
  // the last quad of the cube was written with relative indices: -7 -5 -1 -3 = vertices 6 8 12 10
  const auto &ci = models[1].mesh.indices;
  const auto &cp = models[1].mesh.positions;
  CHECK(cp[3 * ci[30]] == 1.f && cp[3 * ci[31]] == 1.f && cp[3 * ci[32]] == 1.f);  // x = +1 face
  // concatenation for the scene: per-model triples (the 10 line indices give 3 degenerate-looking triples)
  std::vector<float> xyz;
  std::vector<uint32_t> tri;
  detail::concat_models(models, xyz, tri);
  CHECK(xyz.size() == 3 * (4 + 8 + 5 + 6) && tri.size() == 3 * (2 + 12 + 3 + 3));
  CHECK(tri[3 * 17] == 17 && tri[3 * 17 + 1] == 18 && tri[3 * 17 + 2] == 18);
  bool threw = false;
  try {
    tobj::load_obj(tmp + "/missing.obj");
  } catch (const Error &e) {
    threw = e.kind == Error::IOError;
  }
  CHECK(threw);

  // move_to_origin (src/generate.rs:484-527)
  const auto moved = generate::move_to_origin(models);
  float mn[3] = {1e9f, 1e9f, 1e9f};
  for (const auto &m : moved)
    for (size_t i = 0; i < m.mesh.positions.size(); ++i) mn[i % 3] = std::fmin(mn[i % 3], m.mesh.positions[i]);
  CHECK(mn[0] == 0.f && mn[1] == 0.f && mn[2] == 0.f);

  // between_vectors (cgmath): rotates a onto b; identity and half-turn special cases
  for (const Vector3 &a : {Vector3{1, 0, 0}, Vector3{0.6, 0.0, -0.8}, Vector3{0, 0, -1}, Vector3{0, 0, 1}}) {
    const Basis3 R = between_vectors(a, {0, 0, -1});
    const Vector3 Ra{R[0] * a[0] + R[3] * a[1] + R[6] * a[2], R[1] * a[0] + R[4] * a[1] + R[7] * a[2],
                     R[2] * a[0] + R[5] * a[1] + R[8] * a[2]};
    CHECK(dist(Ra, {0, 0, -1}) < 1e-12);
  }

  // path cameras: on the poly-line, looking along +x (the direction of travel is mapped onto -z)
  const auto segs = detail::path_segments(models[3]);
  CHECK(segs.size() == 5);
  detail::Rng rng(5);
  const auto disk = detail::poisson_disk(200, rng);
  CHECK(disk.size() > 60 && disk.size() <= 200);
  const double r = 2.0 * std::sqrt(0.9068996821171089 / (200.0 * 3.141592653589793));
  double closest = 1e9;
  for (size_t i = 0; i < disk.size(); ++i) {
    CHECK(disk[i][0] >= 0.0 && disk[i][0] < 1.0 && disk[i][1] >= 0.0 && disk[i][1] < 1.0);
    for (size_t j = 0; j < i; ++j)
      closest = std::fmin(closest, std::hypot(disk[i][0] - disk[j][0], disk[i][1] - disk[j][1]));
  }
  CHECK(closest >= r);

  std::vector<SnavelyCamera> cams(7, SnavelyCamera::from_position_direction({0, 0, 0}, basis_one()));
  generate::modify_intrinsics(cams, {1.0, -0.1, 0.0}, {2.0, 0.1, 0.0}, 9);
  for (const auto &c : cams) CHECK(c.rec[12] >= 1.0 && c.rec[12] < 2.0 && std::fabs(c.rec[13]) <= 0.1 && c.rec[14] == 0.0);
}

// the graph-editing noise of src/noise.rs:180-378 on a hand-made problem (no GPU needed)
static void graph_noise() {
  std::vector<SnavelyCamera> cams;
  for (int i = 0; i < 4; ++i) cams.push_back(SnavelyCamera::from_position_direction({0.3 * i, 0.0, 0.0}, basis_one()));
  std::vector<Point3> pts;
  for (int i = 0; i < 40; ++i) pts.push_back({-1.0 + 0.05 * i, 0.02 * (i % 7), -3.0 - 0.01 * i});
  VisGraph g(4);
  for (int c = 0; c < 4; ++c)
    for (size_t p = 0; p < pts.size(); ++p) {
      const auto uv = cams[c].project(cams[c].project_world(pts[p]));
      g[c].push_back({p, {uv[0], uv[1]}});
    }
  const BAProblem ba = BAProblem::from_visibility(cams, pts, g);
  const double e0 = ba.total_reprojection_error(2.0);
  CHECK(e0 < 1e-12);

  const BAProblem mis = noise::add_incorrect_correspondences(ba, 0.2, 11);
  CHECK(mis.num_observations() == ba.num_observations() && mis.total_reprojection_error(2.0) > e0);
  for (size_t c = 0; c < 4; ++c) {  // indices are permuted within a camera, (u, v) stay in place
    std::multiset<size_t> a, b;
    for (size_t i = 0; i < g[c].size(); ++i) {
      a.insert(ba.vis_graph[c][i].first);
      b.insert(mis.vis_graph[c][i].first);
      CHECK(ba.vis_graph[c][i].second == mis.vis_graph[c][i].second);
    }
    CHECK(a == b);
  }
  CHECK(noise::add_incorrect_correspondences(ba, 0.0, 11).total_reprojection_error(2.0) == e0);

  const BAProblem dropped = noise::drop_features(ba, 0.25, 12);  // keeps floor(40 * 0.25) per camera
  for (const auto &o : dropped.vis_graph) {
    CHECK(o.size() == 10);
    std::set<size_t> uniq;
    for (const auto &e : o) uniq.insert(e.first);
    CHECK(uniq.size() == 10);
  }
  CHECK(dropped.total_reprojection_error(2.0) < 1e-12);

  const BAProblem split = noise::split_landmarks(ba, 0.5, 13);
  CHECK(split.num_points() == 60 && split.num_observations() == ba.num_observations());
  size_t moved = 0;
  for (const auto &o : split.vis_graph)
    for (const auto &e : o)
      if (e.first >= 40) ++moved;
  CHECK(moved > 10 && moved < 70);                          // about half of the 80 observations of 20 landmarks
  CHECK(split.total_reprojection_error(2.0) < 1e-12);       // copies sit at the same location

  const BAProblem joined = noise::join_landmarks(ba, 0.5, 14);  // 20 observations re-pointed
  size_t changed = 0;
  for (size_t c = 0; c < 4; ++c)
    for (size_t i = 0; i < g[c].size(); ++i)
      if (joined.vis_graph[c][i].first != ba.vis_graph[c][i].first) {
        ++changed;
        // one of the 10 nearest other landmarks: the points are spaced ~0.05 apart along x
        CHECK(dist(pts[joined.vis_graph[c][i].first], pts[ba.vis_graph[c][i].first]) < 0.05 * 10.5);
      }
  CHECK(changed == 20);
  CHECK(joined.total_reprojection_error(2.0) > e0);

  // the k-d tree against brute force
  const noise::detail_n::KdTree tree(pts);
  for (size_t q = 0; q < pts.size(); q += 7) {
    const auto near = tree.nearest(pts[q], 11);
    std::vector<std::pair<double, uint32_t>> all;
    for (size_t i = 0; i < pts.size(); ++i) {
      const Vector3 d = detail::sub(pts[i], pts[q]);
      all.push_back({detail::dot(d, d), (uint32_t)i});
    }
    std::sort(all.begin(), all.end());
    all.resize(11);
    CHECK(near == all);
    CHECK(near[0].second == q);
  }
}

// tests/main.rs:130-201
static void library_properties(const Context &ctx, const std::string &out_path) {
  const BAProblem ba = synthetic::synthetic_grid(ctx, 10, 20, 3, 5., 1., 1., 1., 10., false);
  CHECK(ba.num_cameras() > 0 && ba.num_observations() > 0);
  const double err_start = ba.total_reprojection_error(2.0);
  CHECK(noise::add_drift_normalized(ctx, ba, 0.1, 0.1, 0.1, 1).total_reprojection_error(2.0) > err_start);
  CHECK(noise::add_noise(ctx, ba, 0.1, 0.1, 0.1, 0.1, 2).total_reprojection_error(2.0) > err_start);
  CHECK(noise::add_sin_noise(ctx, ba, {1.0, 1.0, 0.0}, {0.0, 1.0, 0.0}, 1., 2.).total_reprojection_error(2.0) > err_start);
  CHECK(noise::add_drift(ctx, ba, 0.01, 0.0, 0.0, {0.0, 1.0, 0.0}, 3).total_reprojection_error(2.0) > err_start);
  // tests/main.rs:155-190: incorrect_correspondences (>), drop_features (>=), split_landmarks (>=), join_landmarks (>)
  CHECK(noise::add_incorrect_correspondences(ba, 0.01, 4).total_reprojection_error(2.0) > err_start);
  CHECK(noise::drop_features(ba, 0.1, 5).total_reprojection_error(2.0) >= err_start);
  CHECK(noise::split_landmarks(ba, 0.1, 6).total_reprojection_error(2.0) >= err_start);
  CHECK(noise::join_landmarks(ba, 0.01, 7).total_reprojection_error(2.0) > err_start);
  const BAProblem line = synthetic::synthetic_line(ctx, 30, 40, 10., 1., 1., 1., 10., false);
  CHECK(line.num_cameras() > 20);
  // every camera of a culled problem sees > 3 points, every point is seen > 1 times
  std::vector<int> seen(ba.num_points(), 0);
  for (const auto &o : ba.vis_graph) {
    CHECK(o.size() > 3);
    for (size_t i = 0; i < o.size(); ++i) {
      seen[o[i].first]++;
      if (i) CHECK(o[i - 1].first < o[i].first);  // ascending point index (src/generate.rs:446)
    }
  }
  for (int s : seen) CHECK(s > 1);
  std::printf("%s\n", ba.to_string().c_str());
  if (!out_path.empty()) ba.write(out_path);
}

// generate::visibility_graph on a box in front of a camera: the far wall is hidden
static void generate_path(const Context &ctx) {
  // a unit quad wall at z = -2 (two triangles) and points behind / beside it
  const std::vector<float> xyz = {-1, -1, -2, 1, -1, -2, 1, 1, -2, -1, 1, -2};
  const std::vector<uint32_t> tri = {0, 1, 2, 0, 2, 3};
  const Scene scene(ctx, xyz, tri);
  const auto b = scene.bounds();
  CHECK(b.first[2] == -2.0f && b.second[0] == 1.0f);
  const auto hit = scene.intersect({0.f, 0.f, 0.f}, {0.f, 0.f, -1.f});
  CHECK(hit.first && std::fabs(hit.second - 2.0f) < 1e-6f);
  const std::vector<SnavelyCamera> cams = {SnavelyCamera::from_position_direction({0, 0, 0}, basis_one())};
  const std::vector<Point3> pts = {{0.2, 0.1, -4.0} /* behind the wall */, {2.5, 0.0, -4.0} /* beside it */,
                                   {0.0, 0.0, 3.0} /* behind the camera */, {0.1, 0.1, -1.0} /* in front of the wall */};
  const VisGraph g = generate::visibility_graph(scene, cams, pts, 100.0, false);
  CHECK(g.size() == 1 && g[0].size() == 2);
  if (g[0].size() == 2) CHECK(g[0][0].first == 1 && g[0][1].first == 3);
  const auto sampled = generate::generate_world_points_uniform(ctx, xyz, tri, cams, 50, 10.0, 7);
  CHECK(sampled.size() == 50);
  for (const auto &p : sampled) CHECK(p[2] == -2.0 && std::fabs(p[0]) <= 1.0 && std::fabs(p[1]) <= 1.0);
}

// the camera generators of src/generate.rs:109-280 on the small OBJ scene, and the whole
// `generate` sequence of src/bin/city2ba.rs:480-573 through the library calls
static void generate_cameras(const Context &ctx, const std::string &tmp) {
  auto models = tobj::load_obj(write_obj(tmp));
  const tobj::Model path = models[3];
  models.pop_back();
  const Scene scene = generate::commit_scene(ctx, models);
  CHECK(scene.num_triangles() == 2 + 12 + 3);

  const auto on_path = generate::generate_cameras_path(scene, path, 50, 3);
  CHECK(on_path.size() == 50);
  for (const auto &c : on_path) {
    const Point3 p = c.center();
    CHECK(std::fabs(p[1] - 0.5) < 1e-12 && std::fabs(p[2] - 3.0) < 1e-12 && p[0] >= -3.0 - 1e-12 && p[0] <= 2.0 + 1e-12);
    // a point further along the path (+x) is straight ahead: on the -z axis of the camera frame
    const Point3 ahead = c.project_world({p[0] + 1.0, p[1], p[2]});
    CHECK(std::fabs(ahead[0]) < 1e-9 && std::fabs(ahead[1]) < 1e-9 && std::fabs(ahead[2] + 1.0) < 1e-9);
  }
  const auto stepped = generate::generate_cameras_path_step(scene, path, 20, 0.2);
  CHECK(stepped.size() == 20);
  for (size_t i = 0; i < stepped.size(); ++i) CHECK(std::fabs(stepped[i].center()[0] - (-3.0 + 0.2 * i)) < 1e-9);
  bool threw = false;
  try {
    generate::generate_cameras_path_step(scene, path, 100, 0.2);  // 20 > 5: the reference's assert!
  } catch (const std::logic_error &) {
    threw = true;
  }
  CHECK(threw);

  // Poisson cameras: `height` above the tallest surface under them (ground y = 0, cube / roof top y = 2)
  const auto poisson = generate::generate_cameras_poisson(scene, 100, 1.0, 10.0, 21);
  CHECK(poisson.size() > 30 && poisson.size() <= 200);
  size_t on_roof = 0;
  for (const auto &c : poisson) {
    const Point3 p = c.center();
    const bool over_cube = std::fabs(p[0]) < 1.0 && std::fabs(p[2]) < 1.0;
    if (over_cube) ++on_roof;
    if (over_cube || (std::fabs(p[0]) > 1.6 || std::fabs(p[2]) > 1.1)) CHECK(std::fabs(p[1] - (over_cube ? 3.0 : 1.0)) < 1e-5);
    CHECK(p[2] < 0.0 + 10.0);  // the filter of src/generate.rs:264 (z against lower_y + ground)
    CHECK(std::fabs(c.rec[4] - 1.0) < 1e-15);  // yaw about y only
  }
  CHECK(on_roof > 0);
  // ground = 0 keeps only z < lower_y + 0 = 0
  for (const auto &c : generate::generate_cameras_poisson(scene, 100, 1.0, 0.0, 22)) CHECK(c.center()[2] < 0.0);

  // generate -> points -> visibility -> cull, as run_generate does
  auto cams = generate::generate_cameras_path_step(scene, path, 24, 0.2);
  generate::modify_intrinsics(cams, {1, 0, 0}, {1, 0, 0}, 1);
  const auto pts = generate::generate_world_points_uniform(ctx, models, cams, 300, 100.0, 5);
  CHECK(pts.size() == 300);
  const BAProblem ba = BAProblem::from_visibility(cams, pts, generate::visibility_graph(scene, cams, pts, 100.0, false)).cull();
  CHECK(ba.num_cameras() > 0 && ba.num_points() > 0);
  CHECK(ba.total_reprojection_error(1.0) < 1e-9);
}

// test_host obj <file.obj>: one line per model (name, vertices, indices, checksums) for a cross-check against
// the Python mirror's loader on the reference's own OBJ fixtures (tests/test_host_mirror.py)
static int dump_obj(const char *path) {
  for (const auto &m : tobj::load_obj(path)) {
    double ps = 0.0;
    unsigned long long is = 0;
    for (size_t i = 0; i < m.mesh.positions.size(); ++i) ps += (double)m.mesh.positions[i] * (double)(1 + i % 7);
    for (size_t i = 0; i < m.mesh.indices.size(); ++i) is += (unsigned long long)m.mesh.indices[i] * (1 + i % 5);
    std::printf("%s|%zu|%zu|%.9g|%llu\n", m.name.c_str(), m.mesh.positions.size() / 3, m.mesh.indices.size(), ps, is);
  }
  return 0;
}

// malformed / hostile BAL files are ParseErrors (the reference fails in nom / read_exact), never an allocation
// sized by an unvalidated header; empty files parse to empty problems (src/baproblem.rs:580-706)
static void malformed_files(const std::string &tmp) {
  auto write = [&](const std::string &name, const std::string &bytes) {
    const std::string path = tmp + "/" + name;
    FILE *f = std::fopen(path.c_str(), "wb");
    std::fwrite(bytes.data(), 1, bytes.size(), f);
    std::fclose(f);
    return path;
  };
  auto kind_of = [&](const std::string &path) {
    try {
      (void)BAProblem::from_file(path);
    } catch (const Error &e) {
      return (int)e.kind;
    } catch (...) {
      return 99;
    }
    return -1;
  };
  CHECK(kind_of(write("neg.bal", "-1 2 3\n")) == Error::ParseError);
  CHECK(kind_of(write("huge.bal", "99999999999999 1 1\n0 0 0.5 0.5\n")) == Error::ParseError);
  CHECK(kind_of(write("negobs.bal", "1 1 1\n-1 0 0.5 0.5\n")) == Error::ParseError);
  CHECK(kind_of(write("float_index.bal", "1 1 1\n0 1.5 0.5 0.5\n")) == Error::ParseError);
  std::string huge(24, '\xff');  // nc = np = no = 2^64 - 1
  CHECK(kind_of(write("huge.bbal", huge)) == Error::ParseError);
  std::string liar(24, '\0');  // 1 camera, 1 point, 1 observation, then a camera that claims 2^56 observations
  liar[7] = liar[15] = liar[23] = 1;
  liar += std::string("\x01", 1) + std::string(7, '\0');
  liar += std::string(200, '\0');
  CHECK(kind_of(write("liar.bbal", liar)) == Error::ParseError);
  CHECK(kind_of(write("empty.bal", "0 0 0\n")) == -1);  // an empty problem is not an error of from_file
  CHECK(kind_of(write("empty.bbal", std::string(24, '\0'))) == -1);
  CHECK(kind_of(tmp + "/does_not_exist.bal") == Error::IOError);
}

int main(int argc, char **argv) {
  const std::string mode = argc > 1 ? argv[1] : "cpu";
  if (mode == "obj" && argc > 2) return dump_obj(argv[2]);
  const char *tmp = std::getenv("TMPDIR");
  rodrigues_idempotent();
  test_project_world();
  test_project();
  test_project_isomorphic();
  graph_and_io(tmp ? tmp : "/tmp");
  obj_and_cameras(tmp ? tmp : "/tmp");
  graph_noise();
  malformed_files(tmp ? tmp : "/tmp");
  if (mode == "gpu") {
    try {
      const Context ctx(0);
      library_properties(ctx, argc > 2 ? argv[2] : "");
      generate_path(ctx);
      generate_cameras(ctx, tmp ? tmp : "/tmp");
    } catch (const std::exception &e) {
      std::printf("FAILED with exception: %s\n", e.what());
      ++failures;
    }
  }
  std::printf("%s: %d failure(s)\n", mode.c_str(), failures);
  return failures ? 1 : 0;
}
