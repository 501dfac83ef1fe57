// Raw bindings to include/city2ba_cuda.h (ABI version 2).  NOT compiled in this image (no rustc/cargo);
// the same ABI is exercised end to end by city2ba_b200/_lib.py (ctypes) in tests/.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_double, c_float, c_int};

#[repr(C)] pub struct c2b_ctx { _p: [u8; 0] }
#[repr(C)] pub struct c2b_scene { _p: [u8; 0] }

#[repr(C)] #[derive(Clone, Copy)]
pub struct c2b_vis_options {
    pub cull_mode: c_int, pub occlusion: c_int, pub endpoint_guard_rel: c_int,
    pub count_traversal: c_int, pub block_length: c_double, pub block_inset: c_double,
}

#[repr(C)]
pub struct c2b_obs {
    pub n_cameras: u64, pub n_obs: u64,
    pub offsets: *mut u64, pub point_idx: *mut u32, pub uv: *mut c_double,
    pub n_candidates: u64, pub pairs_evaluated: u64, pub nodes_visited: u64, pub tris_tested: u64,
    pub h2d_bytes: u64, pub d2h_bytes: u64,
    pub ms_h2d: c_float, pub ms_prep: c_float, pub ms_cull: c_float, pub ms_sort: c_float,
    pub ms_traverse: c_float, pub ms_compact: c_float, pub ms_d2h: c_float, pub ms_total: c_float,
}

// Embree's RTCRay layout (48 bytes): what embree_rs::Ray wraps (src/generate.rs:253-262, :457-464)
#[repr(C)] #[derive(Clone, Copy)]
pub struct c2b_ray48 {
    pub org_x: c_float, pub org_y: c_float, pub org_z: c_float, pub tnear: c_float,
    pub dir_x: c_float, pub dir_y: c_float, pub dir_z: c_float, pub time: c_float,
    pub tfar: c_float, pub mask: u32, pub id: u32, pub flags: u32,
}

extern "C" {
    pub fn c2b_init(device: c_int, out: *mut *mut c2b_ctx) -> c_int;
    pub fn c2b_shutdown(ctx: *mut c2b_ctx);
    pub fn c2b_last_error() -> *const c_char;
    pub fn c2b_scene_create(ctx: *mut c2b_ctx, xyz: *const c_float, nv: u64,
                            tri: *const u32, nt: u64, out: *mut *mut c2b_scene) -> c_int;
    pub fn c2b_scene_bounds(s: *const c2b_scene, lo: *mut c_float, hi: *mut c_float) -> c_int;
    pub fn c2b_scene_destroy(s: *mut c2b_scene);
    pub fn c2b_intersect1(ctx: *mut c2b_ctx, s: *const c2b_scene, org: *const c_float,
                          dir: *const c_float, hit: *mut c_int, tfar: *mut c_float) -> c_int;
    // closest hit of a whole batch (generate_cameras_poisson's downward rays in one launch): flags = 1 and
    // tfar = distance on a hit; any-hit batch (occluded_stream_aos): tfar = -inf on a hit
    pub fn c2b_intersect(ctx: *mut c2b_ctx, s: *const c2b_scene, rays: *mut c2b_ray48, n: u64) -> c_int;
    pub fn c2b_occluded(ctx: *mut c2b_ctx, s: *const c2b_scene, rays: *mut c2b_ray48, n: u64) -> c_int;
    pub fn c2b_generate_world_points_uniform(ctx: *mut c2b_ctx, xyz: *const c_float, nv: u64, tri: *const u32,
                                             nt: u64, cams: *const c_double, c: u64, num_points: u64,
                                             max_dist: c_double, seed: u64, pts_out: *mut c_double,
                                             n_out: *mut u64) -> c_int;
    pub fn c2b_vis_options_default(opt: *mut c2b_vis_options);
    pub fn c2b_visibility_graph(ctx: *mut c2b_ctx, s: *const c2b_scene, cams: *const c_double,
                                c: u64, pts: *const c_double, p: u64, max_dist: c_double,
                                opt: *const c2b_vis_options, out: *mut c2b_obs) -> c_int;
    pub fn c2b_obs_free(ctx: *mut c2b_ctx, obs: *mut c2b_obs);
    pub fn c2b_add_drift(ctx: *mut c2b_ctx, cams: *mut c_double, c: u64, pts: *mut c_double, p: u64,
                         strength: c_double, angle_strength: c_double, std: c_double,
                         dir: *const c_double, seed: u64) -> c_int;
    pub fn c2b_add_noise(ctx: *mut c2b_ctx, cams: *mut c_double, c: u64, pts: *mut c_double, p: u64,
                         uv: *mut c_double, o: u64, translation_std: c_double, rotation_std: c_double,
                         point_std: c_double, observations_std: c_double, seed: u64) -> c_int;
    pub fn c2b_add_sin_noise(ctx: *mut c2b_ctx, cams: *mut c_double, c: u64, pts: *mut c_double, p: u64,
                             dir: *const c_double, noise_dir: *const c_double, strength: c_double,
                             frequency: c_double) -> c_int;
    // multi-GPU: points that are already on the device (e.g. after an NCCL all-gather)
    pub fn c2b_upload_points_device(ctx: *mut c2b_ctx, d_pts: *const c_double, p: u64) -> c_int;
}
