//! src/gpu_aggregate.rs — the GPU aggregate behind `Scene.aggregate: Arc<dyn Primitive>` (scene.rs:17).
//!
//! `make_aggregate` (renderprocess.rs:1178-1304) walks the `Aggregate.primitives[]` entries, builds one
//! `GeometricPrimitive` per sphere / mesh triangle and pushes it bare or once per `instances[]` transform.  The builder
//! below is fed from the same two arms, in the same order, so that `rrt_hit::prim_id` is the index into the `primitives`
//! vector `BVHAccel::new` would have received.  Shapes are trait objects with private fields there, so the arms hand
//! over what they already hold: the parsed `TriangleMesh` of the obj, the sphere's parameters, the instance transforms.
use std::sync::Arc;

use rrt_sys as ffi;

use crate::geometry::{Bounds3f, IntersectP, Point3f, Ray};
use crate::interaction::SurfaceInteraction;
use crate::primitives::Primitive;
use crate::shape::triangle::TriangleMesh;
use crate::transform::Transform;

fn m16(t: &Transform, inverse: bool) -> [f64; 16] {
    let m = if inverse { &t.m_inv.m } else { &t.m.m };       // transform.rs:177-180: Transform { m, m_inv }
    let mut out = [0.0; 16];
    for r in 0..4 {
        for c in 0..4 {
            out[4 * r + c] = m[r][c];
        }
    }
    out
}
fn flatten(ts: &[Transform], inverse: bool) -> Vec<f64> {
    ts.iter().flat_map(|t| m16(t, inverse).to_vec()).collect()
}
fn to_rrt(r: &Ray) -> ffi::rrt_ray {
    ffi::rrt_ray { o: [r.o.x, r.o.y, r.o.z], d: [r.d.x, r.d.y, r.d.z], t_max: r.t_max, time: r.time }
}

/// Collects the scene while `make_aggregate` runs; `finish` is `BVHAccel::new`.
pub struct GpuAggregateBuilder {
    scene: *mut ffi::rrt_scene,
    prims: Vec<Arc<dyn Primitive>>,
}

impl GpuAggregateBuilder {
    pub fn new(ctx: *mut ffi::rrt_ctx) -> Result<Self, String> {
        let mut scene = std::ptr::null_mut();
        ffi::check(unsafe { ffi::rrt_scene_begin(ctx, &mut scene) })?;
        Ok(Self { scene, prims: Vec::new() })
    }

    /// The "triangle" arm (renderprocess.rs:1229-1282): `gs` = the GeometricPrimitives it made for the obj's triangles,
    /// `instances` = the `make_to_world` of every `instances[]` entry (empty: the bare primitives are pushed).
    /// Mesh vertices are handed over untransformed, as the reference uses them (Q7).
    pub fn add_triangles(&mut self, mesh: &TriangleMesh, material_id: u32, instances: &[Transform],
                         pushed: &[Arc<dyn Primitive>]) -> Result<(), String> {
        let p: Vec<f64> = mesh.p.iter().flat_map(|v| vec![v.x, v.y, v.z]).collect();
        let vi: Vec<u32> = mesh.vertex_indices.iter().map(|&i| i as u32).collect();
        let n: Vec<f64> = mesh.n.iter().flat_map(|v| vec![v.x, v.y, v.z]).collect();
        let ni: Vec<u32> = mesh.normal_indices.iter().map(|&i| i as u32).collect();
        let uv: Vec<f64> = mesh.uv.iter().flat_map(|v| vec![v.x, v.y]).collect();
        let uvi: Vec<u32> = mesh.uv_indices.iter().map(|&i| i as u32).collect();
        let has_n = !n.is_empty() && ni.len() == vi.len();
        let has_uv = !uv.is_empty() && uvi.len() == vi.len();
        let mut mesh_id = 0u32;
        ffi::check(unsafe {
            ffi::rrt_scene_add_mesh(self.scene, mesh.p.len() as u32, p.as_ptr(), mesh.n_triangles as u32, vi.as_ptr(),
                                    if has_n { mesh.n.len() as u32 } else { 0 }, if has_n { n.as_ptr() } else { std::ptr::null() },
                                    if has_n { ni.as_ptr() } else { std::ptr::null() },
                                    if has_uv { mesh.uv.len() as u32 } else { 0 }, if has_uv { uv.as_ptr() } else { std::ptr::null() },
                                    if has_uv { uvi.as_ptr() } else { std::ptr::null() }, &mut mesh_id)
        })?;
        let (m, minv) = (flatten(instances, false), flatten(instances, true));
        ffi::check(unsafe {
            ffi::rrt_scene_add_triangles(self.scene, mesh_id, material_id, instances.len() as u32,
                                         if instances.is_empty() { std::ptr::null() } else { m.as_ptr() },
                                         if instances.is_empty() { std::ptr::null() } else { minv.as_ptr() })
        })?;
        self.prims.extend_from_slice(pushed);   // the same Arcs, in the order make_aggregate pushed them
        Ok(())
    }

    /// The "sphere" arm (renderprocess.rs:1187-1227) over make_sphere (:1097-1106).
    pub fn add_sphere(&mut self, obj_to_world: &Transform, radius: f64, z_min: f64, z_max: f64, phi_max_deg: f64, material_id: u32,
                      instances: &[Transform], pushed: &[Arc<dyn Primitive>]) -> Result<(), String> {
        let (o2w, w2o) = (m16(obj_to_world, false), m16(obj_to_world, true));
        let (m, minv) = (flatten(instances, false), flatten(instances, true));
        ffi::check(unsafe {
            ffi::rrt_scene_add_sphere(self.scene, o2w.as_ptr(), w2o.as_ptr(), radius, z_min, z_max, phi_max_deg, material_id,
                                      instances.len() as u32,
                                      if instances.is_empty() { std::ptr::null() } else { m.as_ptr() },
                                      if instances.is_empty() { std::ptr::null() } else { minv.as_ptr() })
        })?;
        self.prims.extend_from_slice(pushed);
        Ok(())
    }

    /// `BVHAccel::new(primitives, max_prims_in_node, BVHSplitMethod::HLBVH)` (bvh.rs:307-363): flags 0 = the fast tier
    /// (own SAH tree, true closest hit), 1 = the reference's HLBVH with every quirk kept, 2 = tree built on the GPU.
    pub fn finish(self, max_prims_in_node: u32, build_flags: u32) -> Result<GpuAggregate, String> {
        ffi::check(unsafe { ffi::rrt_scene_commit(self.scene, max_prims_in_node, build_flags) })?;
        let mut n = 0u32;
        ffi::check(unsafe { ffi::rrt_scene_num_prims(self.scene, &mut n) })?;
        assert_eq!(n as usize, self.prims.len(), "the GPU scene and the primitive list went out of step");
        let mut b = [0.0f64; 6];
        ffi::check(unsafe { ffi::rrt_world_bound(self.scene, b.as_mut_ptr()) })?;
        Ok(GpuAggregate {
            scene: self.scene,
            prims: self.prims,
            bound: Bounds3f::new(Point3f::new(b[0], b[1], b[2]), Point3f::new(b[3], b[4], b[5])),
        })
    }
}

pub struct GpuAggregate {
    scene: *mut ffi::rrt_scene,
    prims: Vec<Arc<dyn Primitive>>,   // kept for `si.primitive` / the material after a hit
    bound: Bounds3f,
}
// an rrt_scene is immutable after commit and its intersect calls are stream-ordered (rrt.h, "Threading"):
// what `Primitive: Send + Sync` (primitives.rs:14) asks for
unsafe impl Send for GpuAggregate {}
unsafe impl Sync for GpuAggregate {}

impl GpuAggregate {
    pub fn raw(&self) -> *mut ffi::rrt_scene { self.scene }
    /// The batch call: what a tile or a whole frame of rays should use instead of one FFI round trip per ray.
    pub fn intersect_batch(&self, rays: &[ffi::rrt_ray], hits: &mut [ffi::rrt_hit]) -> Result<(), String> {
        assert_eq!(rays.len(), hits.len());
        ffi::check(unsafe { ffi::rrt_intersect(self.scene, rays.len() as u64, rays.as_ptr(), hits.as_mut_ptr()) })
    }
    pub fn intersect_p_batch(&self, rays: &[ffi::rrt_ray], occluded: &mut [u8]) -> Result<(), String> {
        assert_eq!(rays.len(), occluded.len());
        ffi::check(unsafe { ffi::rrt_intersect_p(self.scene, rays.len() as u64, rays.as_ptr(), occluded.as_mut_ptr()) })
    }
}

impl Drop for GpuAggregate {
    fn drop(&mut self) {
        unsafe { ffi::rrt_scene_destroy(self.scene) }
    }
}

impl IntersectP for GpuAggregate {
    fn intersect_p(&self, r: &Ray) -> bool {                       // geometry.rs:94-96, bvh.rs:123-174
        let ray = to_rrt(r);
        let mut occ = 0u8;
        ffi::check(unsafe { ffi::rrt_intersect_p(self.scene, 1, &ray, &mut occ) }).expect("rrt_intersect_p");
        occ != 0
    }
}

impl Primitive for GpuAggregate {
    fn world_bound(&self) -> Bounds3f { self.bound }               // primitives.rs:15, bvh.rs:177-182
    fn intersect(self: Arc<Self>, r: &mut Ray, si: &mut SurfaceInteraction) -> bool {   // bvh.rs:183-236
        let ray = to_rrt(r);
        let mut hit = ffi::rrt_hit::default();
        ffi::check(unsafe { ffi::rrt_intersect(self.scene, 1, &ray, &mut hit) }).expect("rrt_intersect");
        if hit.prim_id == ffi::RRT_NO_HIT {
            return false;
        }
        // The SurfaceInteraction is rebuilt by the reference's own shape code, for the one winner: bound the ray just
        // past the reported distance and let that primitive fill `si` (it also sets r.t_max, primitives.rs:56-57).
        r.t_max = f64::from_bits(hit.t.to_bits() + 1);
        let ok = self.prims[hit.prim_id as usize].clone().intersect(r, si);
        debug_assert!(ok, "the GPU winner must be a hit for the reference's own shape test");
        ok
    }
}
