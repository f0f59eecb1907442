//! src/gpu_integrator.rs — `Box<dyn Integrator>` (integrator/mod.rs:21-23) over the wavefront renderer: the whole of
//! deploy_render's make_scene + make_integrator + render (renderprocess.rs:92-105) in three calls.
use std::ffi::CString;

use rrt_sys as ffi;

use crate::integrator::Integrator;
use crate::renderprocess::write_image;
use crate::geometry::{Bounds2i, Point2i};
use crate::scene::Scene;

pub struct GpuIntegrator {
    scene: *mut ffi::rrt_scene,
    render: *mut ffi::rrt_render,
    xres: i64,
    yres: i64,
    save_to: String,
    /// (tile_mod, tile_rank): this process renders the 16 x 16 sample tiles t with t % tile_mod == tile_rank
    /// (integrator/mod.rs:55-71); (1, 0) = the whole frame
    pub tiles: (u32, u32),
}

impl GpuIntegrator {
    /// `filepath` = the scene.json; `seed` replaces the reference's unseeded thread_rng (Halton permutations, jitter).
    pub fn load(ctx: *mut ffi::rrt_ctx, filepath: &str, save_to: &str, seed: u64, xres: i64, yres: i64) -> Result<Self, String> {
        let path = CString::new(filepath).map_err(|e| e.to_string())?;
        let (mut scene, mut render) = (std::ptr::null_mut(), std::ptr::null_mut());
        ffi::check(unsafe { ffi::rrt_scene_load_json(ctx, path.as_ptr(), std::ptr::null(), seed, &mut scene, &mut render) })?;
        Ok(Self { scene, render, xres, yres, save_to: save_to.to_string(), tiles: (1, 0) })
    }
    pub fn raw(&self) -> *mut ffi::rrt_render { self.render }
}

impl Drop for GpuIntegrator {
    fn drop(&mut self) {
        unsafe {
            ffi::rrt_render_destroy(self.render);
            ffi::rrt_scene_destroy(self.scene);
        }
    }
}

impl Integrator for GpuIntegrator {
    fn render(&mut self, _scene: &Scene) {
        ffi::check(unsafe { ffi::rrt_render_run(self.render, self.tiles.0, self.tiles.1, std::ptr::null()) }).expect("rrt_render_run");
        if self.tiles.1 != 0 {
            return;   // the other ranks hand their tiles to rank 0 (rrt_render_pack_owned + the caller's collective)
        }
        let n = (self.xres * self.yres) as usize;
        let mut rgb = vec![0f64; 3 * n];
        ffi::check(unsafe { ffi::rrt_render_read_film(self.render, rgb.as_mut_ptr(), std::ptr::null_mut()) }).expect("rrt_render_read_film");
        // Film::write_image -> write_image (renderprocess.rs:1501-1530)
        let bounds = Bounds2i::new(Point2i::new(0, 0), Point2i::new(self.xres, self.yres));
        write_image(&self.save_to, &rgb, bounds).expect("write_image");
    }
}
