// lower_impls.rs -- the `lower()` method of every type the reference ships, as the blocks a vecchio maintainer pastes
// into the reference's own modules (the traits' fields are private to them), plus `Camera::lower` and the call in main().
// Companion of rust/gpu.rs (copied to src/gpu.rs) and INTEGRATION.md §2.
//
// STATUS: SOURCE ONLY, like gpu.rs: this image has no rustc / cargo, nothing here has been compiled.  Every block is a
// transliteration of the C++ function of the same name in vecchio_b200/host/vecchio.cpp (`X::lower`), which IS compiled,
// runs in every test, and is compared record by record with an independent restatement of the reference's scene builders
// (tests/test_scene_front.py).  Field names are the reference's (cited per block); record layouts are
// include/vecchio_gpu.h's, mirrored in gpu.rs and checked against the header by tests/test_abi_layouts.py.
//
// Order inside each `lower` matters only for reproducing the C++ front end's record numbering (materials before the
// shape that uses them, a BVH node before its children); the device does not depend on it.

// ======================================================================================================================
// src/hittable.rs -- trait (lines 33-42): two new provided methods
// ======================================================================================================================
//
// pub trait Hittable {
//     fn hit(..); fn bounding_box(..); fn pdf_value(..) {..}; fn random(..) {..}          // unchanged
//
//     /// Flatten into `b`; the default refuses, so a user-defined Hittable fails the upload loudly.
//     fn lower(&self, _b: &mut gpu::Lowering) -> Result<gpu::vk_ref, gpu::GpuError> {
//         Err(gpu::GpuError::Unsupported("Hittable without lower()"))
//     }
//     /// Only `Rect` answers: FlipFace(Rect) -- every FlipFace in src/scene.rs and in Boxy::new -- folds into the rect
//     /// record (VK_RECT_FLIP) instead of a wrapper level.  Trait objects cannot be downcast (no `Any`), hence a method.
//     fn lower_flipped(&self, _b: &mut gpu::Lowering) -> Option<Result<gpu::vk_ref, gpu::GpuError>> { None }
// }

use crate::gpu;

impl Sphere {
    // fields: center, radius, material (src/hittable.rs:47-51)
    fn lower_impl(&self, b: &mut gpu::Lowering) -> Result<gpu::vk_ref, gpu::GpuError> {
        let mat = b.material(&self.material)?;
        b.spheres.push(gpu::vk_sphere { center: gpu::v3(self.center), radius: self.radius }); // radius may be negative (hollow glass)
        b.sphere_mat.push(mat);
        Ok(gpu::vk_mkref(gpu::VK_T_SPHERE, (b.spheres.len() - 1) as u32))
    }
}
// impl Hittable for Sphere { .. fn lower(&self, b: &mut gpu::Lowering) -> Result<gpu::vk_ref, gpu::GpuError> { self.lower_impl(b) } }

impl MovingSphere {
    // fields: center0, center1, time0, time1, radius, material (:137-144)
    fn lower_impl(&self, b: &mut gpu::Lowering) -> Result<gpu::vk_ref, gpu::GpuError> {
        let mat = b.material(&self.material)?;
        b.mspheres.push(gpu::vk_msphere { center0: gpu::v3(self.center0), radius: self.radius, center1: gpu::v3(self.center1),
                                          time0: self.time0, time1: self.time1, mat, _pad: [0; 2] });
        Ok(gpu::vk_mkref(gpu::VK_T_MSPHERE, (b.mspheres.len() - 1) as u32))
    }
}

impl Rect {
    // fields: c0, c1, d0, d1, k, axis0, axis1, axis2, mat (:200-210)
    fn lower_rect(&self, b: &mut gpu::Lowering, flip: bool) -> Result<gpu::vk_ref, gpu::GpuError> {
        let key = self as *const Rect as usize;
        if flip { if let Some(&r) = b.memo_flip.get(&key) { return Ok(r); } }
        let mat = b.material(&self.mat)?;
        let axes = (self.axis0 as u32) | ((self.axis1 as u32) << 2) | ((self.axis2 as u32) << 4) | if flip { gpu::VK_RECT_FLIP } else { 0 };
        b.rects.push(gpu::vk_rect { c0: self.c0, c1: self.c1, d0: self.d0, d1: self.d1, k: self.k, axes, mat, _pad: 0 });
        let r = gpu::vk_mkref(gpu::VK_T_RECT, (b.rects.len() - 1) as u32);
        if flip { b.memo_flip.insert(key, r); } // (the unflipped record is memoised by Lowering::hittable on its Arc)
        Ok(r)
    }
}
// impl Hittable for Rect {
//     fn lower(&self, b: ..) -> Result<..> { self.lower_rect(b, false) }
//     fn lower_flipped(&self, b: ..) -> Option<Result<..>> { Some(self.lower_rect(b, true)) }
// }

impl FlipFace {
    // field: ptr (:295-297)
    fn lower_impl(&self, b: &mut gpu::Lowering) -> Result<gpu::vk_ref, gpu::GpuError> {
        if let Some(r) = self.ptr.lower_flipped(b) { return r; }
        let child = b.hittable(&self.ptr)?;
        Ok(b.push_xform(gpu::VK_X_FLIP, child, 0.0, 0.0, 0.0))
    }
}

// Boxy keeps only box_min / box_max / sides (:314-318): ONE edit to the struct -- remember the material Boxy::new was given
//     pub struct Boxy { box_min: Vec3, box_max: Vec3, sides: Vec<Arc<HittableSS>>, mat: Arc<MaterialSS> }     // + mat
//     Boxy::new(..): `Boxy { box_min: p0, box_max: p1, sides, mat }` with `mat.clone()` for the six sides as today (:325-353)
// The six sides are implied by the record (the device tests the slabs and reports the face index in the order of :325-353).
impl Boxy {
    fn lower_impl(&self, b: &mut gpu::Lowering) -> Result<gpu::vk_ref, gpu::GpuError> {
        let mat = b.material(&self.mat)?;
        b.boxes.push(gpu::vk_box { box_min: gpu::v3(self.box_min), mat, box_max: gpu::v3(self.box_max), _pad: 0 });
        Ok(gpu::vk_mkref(gpu::VK_T_BOX, (b.boxes.len() - 1) as u32))
    }
}

impl ConstantMedium {
    // fields: boundary, phase_function, neg_inv_density (:436-440)
    fn lower_impl(&self, b: &mut gpu::Lowering) -> Result<gpu::vk_ref, gpu::GpuError> {
        let boundary = b.hittable(&self.boundary)?;
        let mat = b.material(&self.phase_function)?;
        b.media.push(gpu::vk_medium { boundary, neg_inv_density: self.neg_inv_density, mat, _pad: 0 });
        Ok(gpu::vk_mkref(gpu::VK_T_MEDIUM, (b.media.len() - 1) as u32))
    }
}

impl Translate {
    // fields: ptr, offset (:501-504)
    fn lower_impl(&self, b: &mut gpu::Lowering) -> Result<gpu::vk_ref, gpu::GpuError> {
        let child = b.hittable(&self.ptr)?;
        Ok(b.push_xform(gpu::VK_X_TRANSLATE, child, self.offset.x, self.offset.y, self.offset.z))
    }
}

// RotateX / RotateY / RotateZ: fields ptr, sin_theta, cos_theta, bb (:631-636, :534-539, :720-725).  The cached box is
// not sent: the device never tests it (the enclosing BVH node's box, which BVHNode::new computed from it, is).
impl RotateX {
    fn lower_impl(&self, b: &mut gpu::Lowering) -> Result<gpu::vk_ref, gpu::GpuError> {
        let child = b.hittable(&self.ptr)?;
        Ok(b.push_xform(gpu::VK_X_ROTATE_X, child, self.sin_theta, self.cos_theta, 0.0))
    }
}
impl RotateY {
    fn lower_impl(&self, b: &mut gpu::Lowering) -> Result<gpu::vk_ref, gpu::GpuError> {
        let child = b.hittable(&self.ptr)?;
        Ok(b.push_xform(gpu::VK_X_ROTATE_Y, child, self.sin_theta, self.cos_theta, 0.0))
    }
}
impl RotateZ {
    fn lower_impl(&self, b: &mut gpu::Lowering) -> Result<gpu::vk_ref, gpu::GpuError> {
        let child = b.hittable(&self.ptr)?;
        Ok(b.push_xform(gpu::VK_X_ROTATE_Z, child, self.sin_theta, self.cos_theta, 0.0))
    }
}

// `impl Hittable for Vec<Arc<HittableSS>>` (:380-434) keeps the default `lower` (unsupported as a WORLD object: no shipped
// scene puts a list in the world; Boxy's `sides` list is implied by vk_box).  The LIGHT list is lowered element by
// element in main() below, which is what its pdf_value / random (:420-433) iterate over.

// ======================================================================================================================
// src/accel.rs -- BVHNode (fields left, right, bb: lines 52-56)
// ======================================================================================================================
impl BVHNode {
    fn lower_impl(&self, b: &mut gpu::Lowering) -> Result<gpu::vk_ref, gpu::GpuError> {
        let i = b.nodes.len(); // depth first, a node before its children, left subtree first
        b.nodes.push(Default::default());
        let left = b.hittable(&self.left)?;
        // `left == right` is the single-object leaf of BVHNode::new (:102-107): the object is tested twice, which matters
        // for a ConstantMedium (two independent free-flight draws) -- the device does the same on equal references
        let right = if Arc::ptr_eq(&self.left, &self.right) { left } else { b.hittable(&self.right)? };
        b.nodes[i] = gpu::vk_node { bb_min: gpu::v3(self.bb.min), left, bb_max: gpu::v3(self.bb.max), right };
        Ok(gpu::vk_mkref(gpu::VK_T_NODE, i as u32))
    }
}

// A user wrapper that only forwards, like `Bowser` (src/scene.rs:340-349, :543-550), forwards `lower` too:
//     impl Hittable for Bowser { .. fn lower(&self, b: &mut gpu::Lowering) -> Result<gpu::vk_ref, gpu::GpuError> { b.hittable(&self.parts) } }

// ======================================================================================================================
// src/material.rs -- traits (lines 20-41, 228-230): one new provided method each
// ======================================================================================================================
//
// pub trait Material { ..; fn lower(&self, _b: &mut gpu::Lowering) -> Result<u32, gpu::GpuError> { Err(gpu::GpuError::Unsupported("Material without lower()")) } }
// pub trait Texture  { ..; fn lower(&self, _b: &mut gpu::Lowering) -> Result<u32, gpu::GpuError> { Err(gpu::GpuError::Unsupported("Texture without lower()")) } }

impl Lambertian {   // field albedo (:46-48)
    fn lower_impl(&self, b: &mut gpu::Lowering) -> Result<u32, gpu::GpuError> {
        let t = b.texture(&self.albedo)?;
        Ok(b.push_material(gpu::VK_M_LAMBERTIAN, t, 0.0, 0))
    }
}
impl Metal {        // fields albedo, fuzz (:112-115)
    fn lower_impl(&self, b: &mut gpu::Lowering) -> Result<u32, gpu::GpuError> {
        let t = b.texture(&self.albedo)?;
        Ok(b.push_material(gpu::VK_M_METAL, t, self.fuzz, 0))
    }
}
impl Dielectric {   // field ref_idx (:145-147)
    fn lower_impl(&self, b: &mut gpu::Lowering) -> Result<u32, gpu::GpuError> {
        Ok(b.push_material(gpu::VK_M_DIELECTRIC, 0, self.ref_idx, 0))
    }
}
impl DiffuseLight { // field emit (:210-212)
    fn lower_impl(&self, b: &mut gpu::Lowering) -> Result<u32, gpu::GpuError> {
        let t = b.texture(&self.emit)?;
        Ok(b.push_material(gpu::VK_M_DIFFUSE_LIGHT, t, 0.0, 0))
    }
}
impl Isotropic {    // field albedo (:437-439)
    fn lower_impl(&self, b: &mut gpu::Lowering) -> Result<u32, gpu::GpuError> {
        let t = b.texture(&self.albedo)?;
        Ok(b.push_material(gpu::VK_M_ISOTROPIC, t, 0.0, 0))
    }
}
impl SpecDiffuse {  // fields specular, diffuse, pct (:468-472); the two children must not be SpecDiffuse themselves:
    // vk_scene_upload refuses a nested one with VK_ERR_UNSUPPORTED (the C++ front end refuses it here)
    fn lower_impl(&self, b: &mut gpu::Lowering) -> Result<u32, gpu::GpuError> {
        let s = b.material(&self.specular)?;
        let d = b.material(&self.diffuse)?;
        Ok(b.push_material(gpu::VK_M_SPECDIFFUSE, s, self.pct, d))
    }
}

impl SolidColor {   // field color_value (:234-236)
    fn lower_impl(&self, b: &mut gpu::Lowering) -> Result<u32, gpu::GpuError> {
        let c = self.color_value;
        Ok(b.push_texture(gpu::VK_TEX_SOLID, [c.x.to_bits(), c.y.to_bits(), c.z.to_bits()]))
    }
}
impl Checker {      // fields odd, even (:245-248)
    fn lower_impl(&self, b: &mut gpu::Lowering) -> Result<u32, gpu::GpuError> {
        let o = b.texture(&self.odd)?;
        let e = b.texture(&self.even)?;
        Ok(b.push_texture(gpu::VK_TEX_CHECKER, [o, e, 0]))
    }
}
impl ImageTexture { // fields buf (RGB8, BPP = 3: :267), width, height (:261-265)
    fn lower_impl(&self, b: &mut gpu::Lowering) -> Result<u32, gpu::GpuError> {
        if b.texels.len() + self.buf.len() > u32::MAX as usize { return Err(gpu::GpuError::Unsupported("texel pool exceeds 4 GiB")); }
        let off = b.texels.len() as u32;
        b.texels.extend_from_slice(&self.buf);
        Ok(b.push_texture(gpu::VK_TEX_IMAGE, [off, self.width as u32, self.height as u32]))
    }
}
impl NoiseTexture { // fields noise: Perlin { random_data, perm_x, perm_y, perm_z } (:306-311), scale (:416-419)
    fn lower_impl(&self, b: &mut gpu::Lowering) -> Result<u32, gpu::GpuError> {
        let mut p = gpu::vk_perlin { ranvec: [[0.0; 3]; 256], perm_x: [0; 256], perm_y: [0; 256], perm_z: [0; 256] };
        for i in 0..256 {
            p.ranvec[i] = gpu::v3(self.noise.random_data[i]);
            p.perm_x[i] = self.noise.perm_x[i] as u8; // permutations of 0..256: fit a byte
            p.perm_y[i] = self.noise.perm_y[i] as u8;
            p.perm_z[i] = self.noise.perm_z[i] as u8;
        }
        b.perlins.push(p);
        let idx = (b.perlins.len() - 1) as u32;
        Ok(b.push_texture(gpu::VK_TEX_NOISE, [idx, self.scale.to_bits(), 0]))
    }
}

// ======================================================================================================================
// src/main.rs -- Camera (fields private to main.rs, lines 57-68) and the call that replaces lines 181-198
// ======================================================================================================================
impl Camera {
    pub fn lower(&self) -> gpu::vk_camera {
        gpu::vk_camera { origin: gpu::v3(self.origin), lower_left_corner: gpu::v3(self.lower_left_corner),
                         horizontal: gpu::v3(self.horizontal), vertical: gpu::v3(self.vertical),
                         u: gpu::v3(self.u), v: gpu::v3(self.v), w: gpu::v3(self.w),
                         lens_radius: self.lens_radius, time0: self.time0, time1: self.time1 }
    }
}

// fn main() -> Result<(), std::io::Error> {                       // error type widened or `.map_err` as the maintainer prefers
//     ...                                                          // scene selection, BVHNode::new: unchanged (:159-169)
//     let world: Arc<HittableSS> = world_bvh.clone();
//     let mut low = gpu::Lowering::default();
//     let root = low.hittable(&world).expect("scene not supported on the GPU path");
//     for l in important.iter() { let r = low.hittable(l).expect("light not supported"); low.lights.push(r); }
//     let mut dev = gpu::Gpu::new(0).expect("no CUDA device");    // or gpu::MultiGpu::new(&[0, 1, 2, 3, 4, 5, 6, 7])
//     dev.upload(&low, root).expect("vk_scene_upload");            // once: the scene stays resident across cameras
//     let seed = 1u64;                                             // new: thread_rng() is unseeded, the GPU path is deterministic per seed
//     for cam in config.cam_iter {
//         let start = Instant::now();
//         let (rgb, _stats) = dev.render(&cam.lower(), width, height, SAMPLES_PER_PIXEL as u32, MAX_DEPTH as u32, seed).expect("vk_render");
//         for (i, pix) in pixels.iter_mut().enumerate() {          // same layout: i = y * width + x, row 0 at the bottom
//             *pix = Vec3::new(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]);
//         }
//         ...                                                      // P3 writer and timing line unchanged (:201-215)
//     }
// }
