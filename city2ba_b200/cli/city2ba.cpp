// city2ba — the reference's command line (src/bin/city2ba.rs) over the C++ host mirror
// (include/city2ba.hpp) and libcity2ba_cuda.so.  Same sub-commands, flags, defaults and stdout lines:
//   generate        src/bin/city2ba.rs:42-113 (options), :480-573 (run_generate)
//   synthetic       :115-151, :447-462
//   synthetic-line  :153-186, :464-478
//   noise           :188-262, :280-357
// (`ply`, :360-445, is a visualisation dump and is not provided.)  The reference draws every random
// number from an unseedable thread_rng(); here `--seed N` (default: from the clock) makes a run
// repeatable.  `--device N` picks the GPU, `--gpus N` (generate, synthetic, synthetic-line) shares the
// visibility graph's cameras between GPUs device .. device+N-1 (c2b_visibility_graph_multi; same output),
// `synthetic --occlusion mesh [--building-height H]` casts rays at the city-block box mesh instead of the
// reference's analytic 2-D wall test.  There is no CPU fallback: without an sm_100 GPU every sub-command
// fails in c2b_init.
//
// Exit status mirrors a Rust binary whose main returns Result<(), city2ba::Error>: 0, 1 with
// `Error: ...` on stderr for an Err, 101 with a panic line for a failed precondition (assert!/panic!/
// unwrap in the reference), 2 for a usage error (structopt / clap).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <map>
#include <set>

#include "city2ba.hpp"

using namespace city2ba;

namespace {

struct UsageError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

// `--name value`, `--name=value`, boolean `--name`, positionals in order
struct Args {
  std::map<std::string, std::string> values;
  std::set<std::string> switches;
  std::vector<std::string> positional;

  Args(int argc, char **argv, int first, const std::map<std::string, std::string> &defaults,
       const std::set<std::string> &bools) {
    values = defaults;
    for (int i = first; i < argc; ++i) {
      std::string a = argv[i];
      if (a.rfind("--", 0) == 0 && a.size() > 2) {
        std::string name = a.substr(2), val;
        const size_t eq = name.find('=');
        const bool has_val = eq != std::string::npos;
        if (has_val) {
          val = name.substr(eq + 1);
          name = name.substr(0, eq);
        }
        if (bools.count(name)) {
          if (has_val) throw UsageError("The argument '--" + name + "' does not take a value");
          switches.insert(name);
        } else if (defaults.count(name)) {
          if (!has_val) {
            if (i + 1 >= argc) throw UsageError("The argument '--" + name + " <" + name + ">' requires a value but none was supplied");
            val = argv[++i];
          }
          values[name] = val;
          given.insert(name);
        } else {
          throw UsageError("Found argument '--" + name + "' which wasn't expected, or isn't valid in this context");
        }
      } else {
        positional.push_back(a);
      }
    }
  }
  std::set<std::string> given;
  bool flag(const std::string &n) const { return switches.count(n) != 0; }
  double f64(const std::string &n) const {
    const std::string &v = values.at(n);
    char *end = nullptr;
    const double d = std::strtod(v.c_str(), &end);
    if (v.empty() || *end) throw UsageError("Invalid value for '--" + n + " <" + n + ">': invalid float literal");
    return d;
  }
  size_t usize(const std::string &n) const {
    const std::string &v = values.at(n);
    char *end = nullptr;
    const unsigned long long d = std::strtoull(v.c_str(), &end, 10);
    if (v.empty() || *end || v[0] == '-') throw UsageError("Invalid value for '--" + n + " <" + n + ">': invalid digit found in string");
    return (size_t)d;
  }
  Vector3 vec3(const std::string &n) const {  // parse_vec3, src/bin/city2ba.rs:25-31
    std::stringstream ss(values.at(n));
    std::string part;
    Vector3 v{};
    for (int k = 0; k < 3; ++k) {
      if (!std::getline(ss, part, ',')) throw std::logic_error("called `Option::unwrap()` on a `None` value");
      char *end = nullptr;
      v[k] = std::strtod(part.c_str(), &end);
      if (part.empty() || *end) throw UsageError("Invalid value for '--" + n + " <" + n + ">': invalid float literal");
    }
    return v;
  }
};

uint64_t seed_of(const Args &a) {
  if (a.given.count("seed")) return (uint64_t)a.usize("seed");
  return (uint64_t)std::chrono::steady_clock::now().time_since_epoch().count() * 0x9E3779B97F4A7C15ull;
}

// --gpus N (not in the reference): GPUs device .. device+N-1 of this box share the visibility graph's cameras
int gpus_of(const Args &a) {
  const size_t n = a.usize("gpus");
  if (n < 1 || n > C2B_MAX_GPUS) throw UsageError("Invalid value for '--gpus <gpus>': expected 1.." + std::to_string(C2B_MAX_GPUS));
  return (int)n;
}

// Rust's `{:.2e}`: two decimals, bare exponent (1.23e4, 5.00e-3, 0.00e0)
std::string sci2(double x) {
  if (std::isnan(x)) return "NaN";
  if (std::isinf(x)) return x > 0 ? "inf" : "-inf";
  char buf[64];
  std::snprintf(buf, sizeof buf, "%.2e", x);
  std::string s(buf);
  const size_t e = s.find('e');
  return s.substr(0, e) + "e" + std::to_string(std::atoi(s.c_str() + e + 1));
}

// ---- city2ba synthetic, src/bin/city2ba.rs:447-462 --------------------------------------------------
int run_synthetic(int argc, char **argv) {
  const Args a(argc, argv, 2,
               {{"cameras-per-block", "10"}, {"points-per-block", "10"}, {"max-dist", "10"}, {"camera-height", "1"},
                {"point-height", "1"}, {"block-inset", "1"}, {"block-length", "20"}, {"blocks", "5"}, {"device", "0"},
                {"seed", "0"}, {"gpus", "1"}, {"occlusion", "analytic"}, {"building-height", "10"}},
               {});
  if (a.positional.size() != 1) throw UsageError("The following required arguments were not provided:\n    <OUTPUT>");
  // every value is parsed (usage errors) before the GPU is touched
  const size_t cpb = a.usize("cameras-per-block"), ppb = a.usize("points-per-block"), blocks = a.usize("blocks");
  const double length = a.f64("block-length"), inset = a.f64("block-inset"), cam_h = a.f64("camera-height"),
               pt_h = a.f64("point-height"), max_dist = a.f64("max-dist");
  const std::string occ = a.values.at("occlusion");
  if (occ != "analytic" && occ != "mesh")
    throw UsageError("Invalid value for '--occlusion <occlusion>': '" + occ + "' (expected analytic or mesh)");
  const double building_h = a.f64("building-height");
  const Context ctx((int)a.usize("device"), gpus_of(a));
  const BAProblem ba = synthetic::synthetic_grid(ctx, cpb, ppb, blocks, length, inset, cam_h, pt_h, max_dist, true,
                                                 occ == "mesh", building_h);
  std::cout << ba.to_string() << "\n";
  ba.write(a.positional[0]);
  return 0;
}

// ---- city2ba synthetic-line, src/bin/city2ba.rs:464-478 -------------------------------------------
int run_synthetic_line(int argc, char **argv) {
  const Args a(argc, argv, 2,
               {{"cameras", "10"}, {"points", "10"}, {"max-dist", "10"}, {"camera-height", "1"}, {"point-height", "1"},
                {"point-offset", "1"}, {"length", "20"}, {"device", "0"}, {"seed", "0"}, {"gpus", "1"}},
               {});
  if (a.positional.size() != 1) throw UsageError("The following required arguments were not provided:\n    <OUTPUT>");
  const size_t n_cams = a.usize("cameras"), n_pts = a.usize("points");
  const double length = a.f64("length"), offset = a.f64("point-offset"), cam_h = a.f64("camera-height"),
               pt_h = a.f64("point-height"), max_dist = a.f64("max-dist");
  const Context ctx((int)a.usize("device"), gpus_of(a));
  const BAProblem ba = synthetic::synthetic_line(ctx, n_cams, n_pts, length, offset, cam_h, pt_h, max_dist, true);
  std::cout << ba.to_string() << "\n";
  ba.write(a.positional[0]);
  return 0;
}

// ---- city2ba noise, src/bin/city2ba.rs:280-357 ------------------------------------------------------
int run_noise(int argc, char **argv) {
  const Args a(argc, argv, 2,
               {{"rotation-std", "0.0"}, {"translation-std", "0.0"}, {"point-std", "0.0"}, {"observation-std", "0.0"},
                {"drift-std", "0.0"}, {"drift-strength", "0.0"}, {"drift-angle", "0.0"}, {"mismatch-chance", "0.0"},
                {"drop-features", "1.0"}, {"split-landmarks", "0.0"}, {"join-landmarks", "0.0"}, {"sin-strength", "0.0"},
                {"sin-frequency", "1.0"}, {"device", "0"}, {"seed", "0"}},
               {"fixed-drift"});
  if (a.positional.size() != 2) throw UsageError("The following required arguments were not provided:\n    <FILE> <OUT>");
  for (const char *k : {"rotation-std", "translation-std", "point-std", "observation-std", "drift-std", "drift-strength",
                        "drift-angle", "mismatch-chance", "drop-features", "split-landmarks", "join-landmarks",
                        "sin-strength", "sin-frequency"})
    (void)a.f64(k);  // usage errors first
  const int device = (int)a.usize("device");
  const uint64_t seed = seed_of(a);
  BAProblem bal = BAProblem::from_file(a.positional[0]);  // src/bin/city2ba.rs:281: a missing file fails before anything else
  const Context ctx(device);
  std::cout << "Initial error: " << sci2(bal.total_reprojection_error(1.)) << " (L1) "
            << sci2(bal.total_reprojection_error(2.)) << " (L2)\n";
  if (a.f64("drop-features") < 1.0) bal = noise::drop_features(bal, a.f64("drop-features"), seed + 1).cull();
  // Join before splitting (src/bin/city2ba.rs:294-298); the reference passes --split-landmarks as the
  // join fraction (:296) and that is kept
  if (a.f64("join-landmarks") > 0.0) bal = noise::join_landmarks(bal, a.f64("split-landmarks"), seed + 2).cull();
  if (a.f64("split-landmarks") > 0.0) bal = noise::split_landmarks(bal, a.f64("split-landmarks"), seed + 3).cull();
  if (a.flag("fixed-drift"))
    bal = noise::add_drift(ctx, bal, a.f64("drift-strength"), a.f64("drift-angle"), a.f64("drift-std"), bal.std(), seed + 4);
  else
    bal = noise::add_drift_normalized(ctx, bal, a.f64("drift-strength"), a.f64("drift-angle"), a.f64("drift-std"), seed + 4);
  if (a.f64("sin-strength") > 0.) {  // sin noise that moves cameras upwards (in positive y)
    bal = noise::add_sin_noise(ctx, bal, {1., 0., 0.}, {0., 1., 0.}, a.f64("sin-strength"), a.f64("sin-frequency"));
    bal = noise::add_sin_noise(ctx, bal, {0., 0., 1.}, {0., 1., 0.}, a.f64("sin-strength"), a.f64("sin-frequency"));
  }
  bal = noise::add_noise(ctx, bal, a.f64("translation-std"), a.f64("rotation-std"), a.f64("point-std"),
                         a.f64("observation-std"), seed + 5);
  bal = noise::add_incorrect_correspondences(bal, a.f64("mismatch-chance"), seed + 6);
  std::cout << "BA Problem with " << bal.num_cameras() << " cameras, " << bal.num_points() << " points, "
            << bal.num_observations() << " correspondences\n";
  std::cout << "Final error: " << sci2(bal.total_reprojection_error(1.)) << " (L1) "
            << sci2(bal.total_reprojection_error(2.)) << " (L2)\n";
  bal.write(a.positional[1]);
  return 0;
}

// ---- city2ba generate, src/bin/city2ba.rs:480-573 ---------------------------------------------------
int run_generate(int argc, char **argv) {
  const Args a(argc, argv, 2,
               {{"cameras", "100"}, {"intrinsics-start", "1,0,0"}, {"intrinsics-end", "1,0,0"}, {"points", "1000"},
                {"max-dist", "100"}, {"ground", "0"}, {"height", "1"}, {"path", ""}, {"step-size", "0"}, {"device", "0"},
                {"seed", "0"}, {"gpus", "1"}, {"predicate", "watertight"}},
               {"no-lcc", "move-to-origin", "compare-predicates"});
  if (a.positional.size() != 2) throw UsageError("The following required arguments were not provided:\n    <FILE> <OUT>");
  if (a.given.count("path") && a.given.count("ground"))
    throw UsageError("The argument '--path <path>' cannot be used with '--ground <ground>'");
  const uint64_t seed = seed_of(a);
  const size_t num_cameras = a.usize("cameras"), num_points = a.usize("points");
  const double max_dist = a.f64("max-dist"), ground = a.f64("ground"), height = a.f64("height"),
               step_size = a.f64("step-size");
  const Vector3 intrinsics_start = a.vec3("intrinsics-start"), intrinsics_end = a.vec3("intrinsics-end");
  const int device = (int)a.usize("device"), gpus = gpus_of(a);
  const std::string pred_name = a.values.at("predicate");
  if (pred_name != "watertight" && pred_name != "mt")
    throw UsageError("Invalid value for '--predicate <predicate>': '" + pred_name + "' (expected watertight or mt)");
  const int predicate = pred_name == "mt" ? C2B_PRED_MT : C2B_PRED_WATERTIGHT;
  std::vector<tobj::Model> models = tobj::load_obj(a.positional[0]);

  std::optional<tobj::Model> model_path;
  if (a.given.count("path")) {
    const std::string &path = a.values.at("path");
    for (const auto &m : models)
      if (m.name == path) {
        model_path = m;
        break;
      }
    if (!model_path) {
      std::string names;
      for (size_t i = 0; i < models.size(); ++i) names += (i ? ", " : "") + models[i].name;
      throw std::logic_error("Could not find a path named " + path + ". Available model names are " + names);
    }
    models.erase(std::remove_if(models.begin(), models.end(), [&](const tobj::Model &m) { return m.name == path; }),
                 models.end());
  }
  if (a.flag("move-to-origin")) models = generate::move_to_origin(std::move(models));

  const Context ctx(device, gpus);
  const Scene cscene = generate::commit_scene(ctx, models);

  std::vector<SnavelyCamera> cameras;
  if (model_path) {
    if (step_size <= 0.0)
      cameras = generate::generate_cameras_path(cscene, *model_path, num_cameras, seed + 1);
    else
      cameras = generate::generate_cameras_path_step(cscene, *model_path, num_cameras, step_size, &std::cout);
  } else {
    cameras = generate::generate_cameras_poisson(cscene, num_cameras, height, ground, seed + 1);
  }
  std::cout << "Generated " << cameras.size() << " cameras\n";

  generate::modify_intrinsics(cameras, intrinsics_start, intrinsics_end, seed + 2);
  std::cout << "Modified intrinsics\n";

  std::vector<Point3> points =
      generate::generate_world_points_uniform(ctx, models, cameras, num_points, max_dist, seed + 3);
  std::cout << "Generated " << points.size() << " world points\n";

  VisGraph vis_graph = generate::visibility_graph(cscene, cameras, points, max_dist, true, predicate);
  size_t edges = 0;
  for (const auto &v : vis_graph) edges += v.size();
  std::cout << "Computed visibility graph with " << edges << " edges\n";
  if (a.flag("compare-predicates")) {
    // how many edges the other occlusion predicate keeps (the two differ on rays that end ON the mesh, which
    // every ray of this generator does: DESIGN.md section 2)
    const int other = predicate == C2B_PRED_MT ? C2B_PRED_WATERTIGHT : C2B_PRED_MT;
    size_t other_edges = 0;
    for (const auto &v : generate::visibility_graph(cscene, cameras, points, max_dist, true, other)) other_edges += v.size();
    std::cout << "Occlusion predicate " << pred_name << ": " << edges << " edges; "
              << (other == C2B_PRED_MT ? "mt" : "watertight") << ": " << other_edges << " edges\n";
  }
  const BAProblem bal = BAProblem::from_visibility(std::move(cameras), std::move(points), std::move(vis_graph));

  // Remove cameras that view too few points and points that are viewed by too few cameras.
  const BAProblem bal_lcc = a.flag("no-lcc") ? bal : bal.cull();
  if (bal_lcc.num_cameras() == 0 || bal_lcc.num_points() == 0) throw Error(Error::EmptyProblem, "No cameras remain");
  std::cout << "Computed LCC with " << bal_lcc.num_cameras() << " cameras, " << bal_lcc.num_points() << " points, "
            << bal_lcc.num_observations() << " edges\n";
  std::cout << "Total reprojection error: " << BAProblem::fmt(bal_lcc.total_reprojection_error(1.)) << "\n";
  bal_lcc.write(a.positional[1]);
  return 0;
}

const char *USAGE =
    "city2ba\nTools for generating synthetic bundle adjustment problems.\n\n"
    "USAGE:\n    city2ba <SUBCOMMAND>\n\n"
    "SUBCOMMANDS:\n"
    "    generate          Generate a synthetic bundle adjustment problem from a 3D model\n"
    "    noise             Add noise to a bundle adjustment problem\n"
    "    synthetic         Generate a synthetic bundle adjustment problem from an grid of city blocks\n"
    "    synthetic-line    Generate a synthetic bundle adjustment problem on a line\n";

}  // namespace

int main(int argc, char **argv) {
  try {
    if (argc < 2) throw UsageError(USAGE);
    const std::string sub = argv[1];
    if (sub == "generate") return run_generate(argc, argv);
    if (sub == "noise") return run_noise(argc, argv);
    if (sub == "synthetic") return run_synthetic(argc, argv);
    if (sub == "synthetic-line") return run_synthetic_line(argc, argv);
    if (sub == "-h" || sub == "--help" || sub == "help") {
      std::cout << USAGE;
      return 0;
    }
    if (sub == "ply") throw UsageError("the `ply` visualisation export is not provided by this build");
    throw UsageError("Found argument '" + sub + "' which wasn't expected, or isn't valid in this context\n\n" + USAGE);
  } catch (const UsageError &e) {
    std::cerr << "error: " << e.what() << "\n";
    return 2;
  } catch (const Error &e) {
    static const char *kinds[] = {"ParseError", "EmptyProblem", "IOError", "Gpu"};
    std::cerr << "Error: " << kinds[e.kind] << "(\"" << e.what() << "\")\n";
    return 1;
  } catch (const std::logic_error &e) {
    std::cerr << "thread 'main' panicked at '" << e.what() << "'\n";
    return 101;
  } catch (const std::bad_alloc &) {
    std::cerr << "Error: IOError(\"out of memory\")\n";
    return 1;
  } catch (const std::exception &e) {
    std::cerr << "Error: IOError(\"" << e.what() << "\")\n";
    return 1;
  }
}
