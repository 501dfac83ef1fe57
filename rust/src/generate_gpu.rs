// The body of generate::visibility_graph over the C ABI (replaces reference src/generate.rs:434-480).
// NOT compiled in this image (no rustc/cargo).
use crate::baproblem::{Camera, SnavelyCamera};
use crate::ffi;
use cgmath::Point3;

/// Owns a c2b_ctx and a committed scene (stands where embree_rs::CommittedScene stood).
pub struct GpuScene {
    pub ctx: *mut ffi::c2b_ctx,
    pub raw: *mut ffi::c2b_scene,
}

impl GpuScene {
    /// Scene::new + model_to_geometry + attach_geometry + commit (src/bin/city2ba.rs:515-521)
    pub fn new(models: &[tobj::Model]) -> GpuScene {
        let (mut xyz, mut tri, mut base) = (Vec::<f32>::new(), Vec::<u32>::new(), 0u32);
        for m in models {
            xyz.extend_from_slice(&m.mesh.positions);
            // model_to_geometry regroups EACH model's index list into triples on its own
            // (num_tri = indices.len() / 3, src/generate.rs:78): drop a model's trailing 1-2 indices (an
            // odd number of `l` pairs) here, or every later model's triangles would shift
            let n = m.mesh.indices.len() / 3 * 3;
            tri.extend(m.mesh.indices[..n].iter().map(|i| i + base));
            base += (m.mesh.positions.len() / 3) as u32;
        }
        let mut ctx = std::ptr::null_mut();
        let mut raw = std::ptr::null_mut();
        unsafe {
            assert!(ffi::c2b_init(0, &mut ctx) == 0);
            assert!(ffi::c2b_scene_create(ctx, xyz.as_ptr(), (xyz.len() / 3) as u64, tri.as_ptr(),
                                          (tri.len() / 3) as u64, &mut raw) == 0);
        }
        GpuScene { ctx, raw }
    }
}

impl Drop for GpuScene {
    fn drop(&mut self) {
        unsafe {
            ffi::c2b_scene_destroy(self.raw);
            ffi::c2b_shutdown(self.ctx);
        }
    }
}

pub fn visibility_graph(scene: &GpuScene, cameras: &[SnavelyCamera], points: &[Point3<f64>],
                        max_dist: f64, _verbose: bool) -> Vec<Vec<(usize, (f64, f64))>> {
    // marshal: SnavelyCamera has no repr(C) (src/baproblem.rs:130-138) -> 15 doubles per camera:
    // dir as column-major 3x3 (cgmath Matrix3 is [[S;3];3] by column), loc, intrin
    let mut cams = Vec::with_capacity(cameras.len() * 15);
    for c in cameras {
        let m: &[[f64; 3]; 3] = c.dir.as_ref().as_ref();
        for col in m { cams.extend_from_slice(col); }
        cams.extend_from_slice(&[c.loc.x, c.loc.y, c.loc.z, c.intrin.x, c.intrin.y, c.intrin.z]);
    }
    // Point3<f64> is repr(C): the slice is already P x 3 doubles
    let pts = points.as_ptr() as *const f64;
    let mut opt = unsafe { std::mem::zeroed() };
    let mut out: ffi::c2b_obs = unsafe { std::mem::zeroed() };
    unsafe {
        ffi::c2b_vis_options_default(&mut opt);
        let rc = ffi::c2b_visibility_graph(scene.ctx, scene.raw, cams.as_ptr(), cameras.len() as u64,
                                           pts, points.len() as u64, max_dist, &opt, &mut out);
        assert!(rc == 0, "{}", std::ffi::CStr::from_ptr(ffi::c2b_last_error()).to_string_lossy());
        let off = std::slice::from_raw_parts(out.offsets, cameras.len() + 1);
        let idx = std::slice::from_raw_parts(out.point_idx, out.n_obs as usize);
        let uv = std::slice::from_raw_parts(out.uv, 2 * out.n_obs as usize);
        let graph = (0..cameras.len()).map(|c| {
            (off[c] as usize..off[c + 1] as usize)
                .map(|i| (idx[i] as usize, (uv[2 * i], uv[2 * i + 1]))).collect()
        }).collect();
        ffi::c2b_obs_free(scene.ctx, &mut out);
        graph
    }
}
