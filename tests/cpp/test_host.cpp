// tests/cpp/test_host.cpp — the reference's own tests restated against the C++ host mirror
// (include/city2ba.hpp).  Mirrors: the inline unit tests of src/baproblem.rs:64-75,227-249 and the
// library property tests of tests/main.rs:130-201 (every noise function must increase the total
// reprojection error of synthetic_grid(10,20,3,5,1,1,1,10); synthetic_line keeps > 20 cameras),
// plus BAL text / binary round trips (src/baproblem.rs:580-785).
//   test_host cpu            camera math, BAL I/O and cull on hand-made data (no GPU)
//   test_host gpu <out.bbal> everything; writes the culled synthetic grid for a cross-check in pytest
#include <cstdio>
#include <cstdlib>
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

// tests/main.rs:130-201
static void library_properties(const Context &ctx, const std::string &out_path) {
  const BAProblem ba = synthetic::synthetic_grid(ctx, 10, 20, 3, 5., 1., 1., 1., 10., false);
  CHECK(ba.num_cameras() > 0 && ba.num_observations() > 0);
  const double err_start = ba.total_reprojection_error(2.0);
  CHECK(noise::add_drift_normalized(ctx, ba, 0.1, 0.1, 0.1, 1).total_reprojection_error(2.0) > err_start);
  CHECK(noise::add_noise(ctx, ba, 0.1, 0.1, 0.1, 0.1, 2).total_reprojection_error(2.0) > err_start);
  CHECK(noise::add_sin_noise(ctx, ba, {1.0, 1.0, 0.0}, {0.0, 1.0, 0.0}, 1., 2.).total_reprojection_error(2.0) > err_start);
  CHECK(noise::add_drift(ctx, ba, 0.01, 0.0, 0.0, {0.0, 1.0, 0.0}, 3).total_reprojection_error(2.0) > err_start);
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

int main(int argc, char **argv) {
  const std::string mode = argc > 1 ? argv[1] : "cpu";
  const char *tmp = std::getenv("TMPDIR");
  rodrigues_idempotent();
  test_project_world();
  test_project();
  test_project_isomorphic();
  graph_and_io(tmp ? tmp : "/tmp");
  if (mode == "gpu") {
    try {
      const Context ctx(0);
      library_properties(ctx, argc > 2 ? argv[2] : "");
      generate_path(ctx);
    } catch (const std::exception &e) {
      std::printf("FAILED with exception: %s\n", e.what());
      ++failures;
    }
  }
  std::printf("%s: %d failure(s)\n", mode.c_str(), failures);
  return failures ? 1 : 0;
}
