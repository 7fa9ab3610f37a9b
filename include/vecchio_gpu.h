/*
 * vecchio_gpu.h -- C ABI of the B200 (sm_100a) path-tracing sample loop.
 *
 * This is the drop-in boundary for the ONE hot path of browserdotsys/vecchio: the
 * per-pixel sample loop of the reference, `src/main.rs:181-198`, and everything it
 * calls (`ray_color` `src/main.rs:123-153`, every `Hittable::hit`, `Material::
 * scatter_with_pdf`, `Texture::value`, the PDF objects of `src/util.rs`).
 *
 * The reference has no FFI of its own (pure safe Rust).  The entry points below are
 * what a `gpu` module of the Rust host would bind with `extern "C"` (see
 * INTEGRATION.md and rust/gpu.rs); each one names the reference code it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; caller owns every host buffer; the library
 *     keeps no pointer after a call returns.
 *   - every function returns VK_OK (0) or a negative vk_status; the message is
 *     available from vk_last_error().  Nothing aborts or throws across the ABI.
 *   - there is NO CPU fallback: without a CUDA device vk_create fails with
 *     VK_ERR_NO_DEVICE, and a scene element the GPU path does not implement fails
 *     vk_scene_upload with VK_ERR_UNSUPPORTED.
 *   - all arithmetic is IEEE fp32, like the reference (`f32`, `src/vec3.rs:3-8`).
 */
#ifndef VECCHIO_GPU_H
#define VECCHIO_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VK_API_VERSION 1u

typedef enum vk_status {
    VK_OK = 0,
    VK_ERR_INVALID = -1,     /* bad argument / malformed scene description          */
    VK_ERR_NO_DEVICE = -2,   /* no CUDA device (there is no CPU fallback)           */
    VK_ERR_CUDA = -3,        /* a CUDA runtime call failed                          */
    VK_ERR_UNSUPPORTED = -4, /* scene element not implemented on the GPU path       */
    VK_ERR_NO_SCENE = -5,    /* render/intersect before vk_scene_upload             */
    VK_ERR_OOM = -6
} vk_status;

/* ------------------------------------------------------------------------------
 * Flattened scene.  The reference's object graph of `Arc<dyn Hittable>` trait
 * objects (`src/hittable.rs:33-44`) is lowered to one typed record array per
 * primitive kind plus 32-bit tagged references between them.
 * ---------------------------------------------------------------------------- */

typedef uint32_t vk_ref; /* (type << 28) | index ; 0 == none */

enum {
    VK_T_NONE = 0,
    VK_T_NODE = 1,    /* BVHNode            src/accel.rs:52-83                 */
    VK_T_SPHERE = 2,  /* Sphere             src/hittable.rs:46-121             */
    VK_T_MSPHERE = 3, /* MovingSphere       src/hittable.rs:136-197            */
    VK_T_RECT = 4,    /* Rect (+FlipFace)   src/hittable.rs:199-312            */
    VK_T_BOX = 5,     /* Boxy               src/hittable.rs:314-378            */
    VK_T_XFORM = 6,   /* Translate/Rotate{X,Y,Z}/FlipFace  :294-312, :500-807  */
    VK_T_MEDIUM = 7   /* ConstantMedium     src/hittable.rs:436-498            */
};
#define VK_REF_NONE 0u
#define VK_REF(type, index) ((((uint32_t)(type)) << 28) | ((uint32_t)(index) & 0x0FFFFFFFu))
#define VK_REF_TYPE(r) (((uint32_t)(r)) >> 28)
#define VK_REF_INDEX(r) (((uint32_t)(r)) & 0x0FFFFFFFu)

/* BVHNode {left, right, bb} (src/accel.rs:52-56).  `left == right` marks the
 * single-object leaf of BVHNode::new (src/accel.rs:102-107): the object is tested
 * twice, which matters for ConstantMedium (two independent free-flight draws). */
typedef struct vk_node {
    float bb_min[3];
    vk_ref left;
    float bb_max[3];
    vk_ref right;
} vk_node; /* 32 B, two 128-bit loads */

typedef struct vk_sphere {
    float center[3];
    float radius; /* may be negative (hollow glass, src/scene.rs:123-127) */
} vk_sphere;      /* 16 B; material in vk_scene_desc.sphere_mat[] */

typedef struct vk_msphere {
    float center0[3];
    float radius;
    float center1[3];
    float time0;
    float time1;
    uint32_t mat;
    uint32_t _pad[2];
} vk_msphere; /* 48 B */

#define VK_RECT_FLIP 0x100u
typedef struct vk_rect {
    float c0, c1, d0, d1; /* bounds on axis0 / axis1 (src/hittable.rs:200-210) */
    float k;              /* plane position on axis2                            */
    uint32_t axes;        /* axis0 | axis1<<2 | axis2<<4 | VK_RECT_FLIP         */
    uint32_t mat;
    uint32_t _pad;
} vk_rect; /* 32 B */

typedef struct vk_box {
    float box_min[3];
    uint32_t mat;
    float box_max[3];
    uint32_t _pad;
} vk_box; /* 32 B; the six sides of Boxy::new (src/hittable.rs:325-353) are implied */

enum {
    VK_X_TRANSLATE = 0, /* a,b,c = offset          src/hittable.rs:500-532 */
    VK_X_ROTATE_X = 1,  /* a = sin, b = cos        src/hittable.rs:631-718 */
    VK_X_ROTATE_Y = 2,  /*                         src/hittable.rs:534-629 */
    VK_X_ROTATE_Z = 3,  /*                         src/hittable.rs:720-807 */
    VK_X_FLIP = 4       /* FlipFace of a non-Rect  src/hittable.rs:294-312 */
};
#define VK_MAX_XFORM_DEPTH 8
typedef struct vk_xform {
    uint32_t kind;
    vk_ref child;
    uint32_t _pad0[2];
    float a, b, c;
    uint32_t _pad1;
} vk_xform; /* 32 B; one record per wrapper level, chains are walked child by child */

typedef struct vk_medium {
    vk_ref boundary;
    float neg_inv_density; /* -1/density, src/hittable.rs:447 */
    uint32_t mat;          /* the Isotropic phase material      */
    uint32_t _pad;
} vk_medium; /* 16 B */

enum {
    VK_M_LAMBERTIAN = 0,    /* src/material.rs:45-109  */
    VK_M_METAL = 1,         /* src/material.rs:111-142 */
    VK_M_DIELECTRIC = 2,    /* src/material.rs:144-207 */
    VK_M_DIFFUSE_LIGHT = 3, /* src/material.rs:209-226 */
    VK_M_ISOTROPIC = 4,     /* src/material.rs:436-465 */
    VK_M_SPECDIFFUSE = 5    /* src/material.rs:467-488 */
};
typedef struct vk_material {
    uint32_t type;
    uint32_t tex;   /* albedo / emit texture index; SPECDIFFUSE: specular material index */
    float param;    /* METAL fuzz | DIELECTRIC ref_idx | SPECDIFFUSE pct                 */
    uint32_t aux;   /* SPECDIFFUSE: diffuse material index                               */
} vk_material;      /* 16 B */

enum {
    VK_TEX_SOLID = 0,   /* src/material.rs:233-242 */
    VK_TEX_CHECKER = 1, /* src/material.rs:244-259 */
    VK_TEX_IMAGE = 2,   /* src/material.rs:261-304 */
    VK_TEX_NOISE = 3    /* src/material.rs:416-434 */
};
typedef struct vk_texture {
    uint32_t type;
    union {
        float rgb[3];                                                  /* SOLID   */
        struct { uint32_t odd, even, _p; } checker;                    /* CHECKER */
        struct { uint32_t texel_offset, width, height; } image;        /* IMAGE: RGB8, BPP=3 */
        struct { uint32_t perlin; float scale; uint32_t _p; } noise;   /* NOISE   */
    };
} vk_texture; /* 16 B */

/* Perlin tables (src/material.rs:306-311), generated on the host (:357-377). */
typedef struct vk_perlin {
    float ranvec[256][3];
    uint8_t perm_x[256];
    uint8_t perm_y[256];
    uint8_t perm_z[256];
} vk_perlin;

typedef struct vk_scene_desc {
    uint32_t api_version; /* VK_API_VERSION */
    vk_ref root;          /* the world: `BVHNode::new(&mut config.world)` src/main.rs:168 */

    const vk_node* nodes;        uint32_t n_nodes;
    const vk_sphere* spheres;    const uint32_t* sphere_mat; uint32_t n_spheres;
    const vk_msphere* mspheres;  uint32_t n_mspheres;
    const vk_rect* rects;        uint32_t n_rects;
    const vk_box* boxes;         uint32_t n_boxes;
    const vk_xform* xforms;      uint32_t n_xforms;
    const vk_medium* media;      uint32_t n_media;

    const vk_ref* lights;        uint32_t n_lights; /* SceneConfig.lights src/scene.rs:19 */

    const vk_material* materials; uint32_t n_materials;
    const vk_texture* textures;   uint32_t n_textures;
    const uint8_t* texels;        uint64_t n_texel_bytes;
    const vk_perlin* perlins;     uint32_t n_perlins;
} vk_scene_desc;

/* Camera (src/main.rs:56-68): the ten fields, in declaration order. */
typedef struct vk_camera {
    float origin[3];
    float lower_left_corner[3];
    float horizontal[3];
    float vertical[3];
    float u[3], v[3], w[3];
    float lens_radius;
    float time0, time1;
} vk_camera; /* 24 floats */

/* AUTO picks by the evidence recorded in DESIGN.md / profiles/.  MEGAKERNEL = one lane, one path
 * (persistent threads, path regeneration); WAVEFRONT = generate / extend / shade as separate kernels
 * over a slot pool in device memory; STAGED = the persistent megakernel with the wavefront's stages
 * and material sort inside each CTA (slot pool in shared memory); WARPQ = the same stages inside each
 * WARP: per-warp slot pool and per-class queues in shared memory, no barrier anywhere; STEPQ = warp
 * queues for BVH scenes with the traversal itself cut into queued steps (node visit / sphere / box /
 * other leaves), traversal state in shared memory (a flat-program scene runs WARPQ). */
enum { VK_VARIANT_AUTO = 0, VK_VARIANT_MEGAKERNEL = 1, VK_VARIANT_WAVEFRONT = 2, VK_VARIANT_STAGED = 3, VK_VARIANT_WARPQ = 4,
       VK_VARIANT_STEPQ = 5 };
enum {
    VK_FLAG_STRICT_MATH = 1u, /* no FMA contraction, IEEE div/sqrt, division slab test:
                                 the op sequence of the reference, for hit parity    */
    VK_FLAG_FORCE_BVH = 2u,   /* traverse the BVH even when the scene is small enough for
                                 the flat (divergence-free) traversal program         */
    VK_FLAG_LEGACY_SCATTER = 4u, /* the book-1/2 integrator of the reference's legacy `Material::scatter`
                                 methods (src/material.rs:21-28, 85-90, 118-132, 150-175, 215-217,
                                 442-446): no light list, no PDFs, `emitted + attenuation *
                                 ray_color(scattered)`.  HEAD itself cannot render a scene without
                                 lights (`choose().unwrap()` panics, src/hittable.rs:431)          */
    VK_FLAG_SKY_BACKGROUND = 8u  /* a missed ray returns the book-1 sky (1-t)*white + t*(0.5,0.7,1),
                                 t = 0.5*(unit(d).y + 1) (sample/inoneweekend.png) instead of
                                 `background` (src/main.rs:124)                                    */
};

/* The constants of src/main.rs:28-29,171-172 and the choices the reference leaves
 * to its (unseeded) RNG, as parameters. */
typedef struct vk_render_params {
    uint32_t width, height;
    uint32_t spp;       /* SAMPLES_PER_PIXEL: the divisor of the mean (main.rs:196)     */
    uint32_t spp_begin; /* this call renders global samples [spp_begin, spp_begin+spp_count) */
    uint32_t spp_count; /* 0 == all of them (spp - spp_begin)                            */
    uint32_t max_depth; /* MAX_DEPTH, `depth > MAX_DEPTH` semantics (main.rs:126)        */
    uint64_t seed;      /* Philox4x32-10 key                                             */
    float background[3];/* src/main.rs:124 (always 0 at HEAD)                            */
    uint32_t variant;   /* VK_VARIANT_*                                                  */
    uint32_t flags;     /* VK_FLAG_*                                                     */
} vk_render_params;

typedef struct vk_stats {
    uint64_t paths;           /* samples started = W*H*spp_count                         */
    uint64_t rays;            /* world.hit() segment queries from ray_color (main.rs:130)*/
    uint64_t dropped_samples; /* non-finite samples filtered (main.rs:191-194)           */
    float ms_kernels;         /* CUDA-event time of the render kernels                   */
    float ms_total;           /* CUDA-event time incl. D2H of the image (vk_render)      */
    uint32_t variant;         /* variant that ran                                        */
    uint32_t launches;        /* kernels launched by this call                           */
    uint64_t node_visits;     /* BVH nodes fetched (wide nodes: one fetch tests two boxes) */
    uint64_t prim_tests;      /* primitive / instance / medium tests                      */
} vk_stats;

typedef struct vk_ray {
    float origin[3];
    float direction[3];
    float time;
    float tmin, tmax;
} vk_ray; /* 36 B */

/* HitRec (src/hittable.rs:11-20) plus the primitive id the reference lacks. */
typedef struct vk_hit {
    vk_ref prim;    /* leaf record that produced the hit; VK_REF_NONE == miss */
    uint32_t face;  /* BOX: side index 0..5 in Boxy::new order; else 0        */
    uint32_t mat;
    uint32_t front;
    float t;
    float p[3];
    float normal[3];
    float u, v;
    uint32_t _pad;
} vk_hit; /* 56 B */

#define VK_MEDIUM_XI_SLOTS 8 /* vk_intersect's injected table: slot = medium index * 2 + second-visit; scenes of up to 4 media */

typedef struct vk_ctx vk_ctx;

/* Create a context on CUDA device `device`.  Replaces nothing in the reference
 * (it has no device); owns the stream, device scene and accumulation buffers. */
int vk_create(int device, vk_ctx** out);
void vk_destroy(vk_ctx* ctx);
/* Message of the last failure on `ctx` (or of the last failed vk_create if NULL). */
const char* vk_last_error(const vk_ctx* ctx);

/* Copy a flattened scene to the device (host arrays are borrowed for the call).
 * Replaces the `Arc<BVHNode>` world / `Arc<Vec<..>>` lights handed to the loop at
 * src/main.rs:168-169. */
int vk_scene_upload(vk_ctx* ctx, const vk_scene_desc* scene);

/* Host-only: validate a flattened scene and plan its device layout exactly as vk_scene_upload would,
 * without a device.  Reports what the upload would build (and lets CPU tests cover the validator, the
 * 4-wide BVH collapse and the flat-program builder).  Same return codes as vk_scene_upload; the message
 * goes to err (nullable). */
typedef struct vk_scene_info {
    uint32_t flat_entries;         /* primitives in the flat traversal program, 0 = the BVH is traversed   */
    uint32_t flat_segments;        /* world frame + instance frames                                        */
    uint32_t flat_subtrees;        /* hybrid program: homogeneous subtrees kept as BVH entries (0 = pure)  */
    uint32_t simple;               /* 1 = the trimmed ("simple scene") build of the staged kernel applies  */
    uint32_t wide_nodes;           /* 4-wide nodes made from the reference's binary nodes                  */
    uint32_t wide_levels_world;    /* 4-wide levels of the world BVH / of the deepest instanced sub-BVH     */
    uint32_t wide_levels_instance;
    uint32_t stack_need;           /* traversal stack entries needed (<= 96)                               */
    uint32_t dynamic_megakernel;   /* 1 = BVH large enough for the dynamic re-fill megakernel               */
    uint32_t flat_boxes;           /* Boxy entries of the flat program (render build: one slab test each)   */
    uint32_t flat_direct;          /* hit-table entries whose HitRec the shade stage writes down directly   */
} vk_scene_info;
int vk_scene_check(const vk_scene_desc* scene, vk_scene_info* info, char* err, size_t err_len);

/* The sample loop, src/main.rs:181-198.  out_rgb: W*H*3 floats, pixel i = y*W+x,
 * row 0 = bottom (main.rs:182-183), linear mean over `spp` (main.rs:196).
 * out_sumsq (nullable): per channel sum of squares of the kept samples.  Host buffers. */
int vk_render(vk_ctx* ctx, const vk_camera* cam, const vk_render_params* params,
              float* out_rgb, float* out_sumsq, vk_stats* stats);

/* The sample loop followed by the output conversion of src/main.rs:201-214 on the device: every
 * channel goes through Vec3::to_color (src/vec3.rs:54-61: (256 * clamp(sqrt(c), 0, 0.999)) as u32, NaN
 * -> 0) and the rows are written top-down (y = height-1 first, main.rs:209), i.e. out_rgb8 holds the
 * W*H*3 numbers of the reference's P3 file in file order.  Only a quarter of the bytes cross PCIe.
 * A turntable (RotatingCamera, src/scene.rs:65-91) calls this once per camera with the scene resident. */
int vk_render_rgb8(vk_ctx* ctx, const vk_camera* cam, const vk_render_params* params,
                   uint8_t* out_rgb8, vk_stats* stats);

/* Same loop, device-resident result: writes per-pixel SUMS (not means) of samples
 * [spp_begin, spp_begin+spp_count) to d_sum (and d_sumsq, nullable), both device
 * pointers of W*H*3 floats, enqueued on the context stream; synchronised before return
 * only when stats != NULL.  This is the per-GPU spp slice; slices are combined by one NCCL
 * reduce. */
int vk_render_device(vk_ctx* ctx, const vk_camera* cam, const vk_render_params* params,
                     float* d_sum, float* d_sumsq, vk_stats* stats);

/* d_rgb[i] = d_sum[i] / spp on the device (main.rs:196), after the reduce.  Asynchronous on the
 * context stream. */
int vk_finalize_device(vk_ctx* ctx, const float* d_sum, float* d_rgb, size_t n_floats,
                       uint32_t spp);

/* Enqueue all following work on `stream` (a cudaStream_t of the caller, e.g. the one its NCCL
 * reduce runs on); NULL restores the context's own stream. */
int vk_set_stream(vk_ctx* ctx, void* stream);

/* Counters accumulated since the last read (paths, rays, dropped samples, kernel launches);
 * synchronises the stream.  vk_render_device with stats == NULL is fully asynchronous and leaves
 * its counts to be collected here. */
int vk_flush_stats(vk_ctx* ctx, vk_stats* stats);

/* One context over several GPUs of a node: what replaces the rayon pool of src/main.rs:181 when the host has more
 * than one device.  The frame's samples are split by GLOBAL sample index (device k of N renders
 * [k * count / N, (k + 1) * count / N) of the requested range, for every pixel), every device keeps its slice in
 * 64-bit fixed-point accumulators, and device devices[0] adds its peers' accumulators to its own with a kernel that
 * reads them over NVLink (peer-mapped memory; no communicator, no host staging).  Integer addition: the frame is
 * bit-identical to the frame one GPU renders for the same seed, whatever N.  Same conventions as the single-device
 * calls; vk_multi_last_error(NULL) reports a failed vk_multi_create. */
typedef struct vk_multi vk_multi;
int vk_multi_create(const int* devices, int n_devices, vk_multi** out);
void vk_multi_destroy(vk_multi* m);
const char* vk_multi_last_error(const vk_multi* m);
int vk_multi_device_count(const vk_multi* m);
int vk_multi_scene_upload(vk_multi* m, const vk_scene_desc* scene);            /* the scene is replicated on every device */
int vk_multi_render(vk_multi* m, const vk_camera* cam, const vk_render_params* params, float* out_rgb, float* out_sumsq,
                    vk_stats* stats);                                           /* vk_render over all devices              */
int vk_multi_render_rgb8(vk_multi* m, const vk_camera* cam, const vk_render_params* params, uint8_t* out_rgb8,
                         vk_stats* stats);                                      /* vk_render_rgb8 over all devices         */

/* Parity hook: closest hit of `world.hit(&r, tmin, tmax)` (src/accel.rs:58-83) for
 * a batch of rays.  medium_xi (nullable): n * VK_MEDIUM_XI_SLOTS uniform variates
 * for ConstantMedium::hit's free-flight draw (src/hittable.rs:473). */
int vk_intersect(vk_ctx* ctx, const vk_ray* rays, size_t n, const float* medium_xi,
                 uint32_t flags, vk_hit* out);

/* Parity hook for the SHADING half of the path, next to vk_intersect for the geometry half: one call of the
 * device functions behind `Material::scatter_with_pdf` / `scattering_pdf` / `emitted` (src/material.rs:92-108,
 * 134-141, 177-206, 218-225, 448-487), `Texture::value` (:238-303, 430-433), the light list's `pdf_value` /
 * `random` (src/hittable.rs:104-134, 271-291, 371-377, 420-433) and the mixture estimator that ties them
 * together in `ray_color` (src/main.rs:131-149) -- on the uploaded scene's materials, textures and lights, with
 * the variates SUPPLIED so that the oracle can be fed the same ones.
 *   VK_EVAL_BOUNCE         one pass of ray_color's body after world.hit() returned the given HitRec:
 *                          beta starts at (1,1,1), L at 0; out = continue?, scattered ray, weight, emission
 *   VK_EVAL_BOUNCE_LEGACY  the same for the legacy `Material::scatter` integrator
 *   VK_EVAL_TEXTURE        textures[index].value(u, v, p) -> beta
 *   VK_EVAL_LIGHTS_PDF     lights.pdf_value(p, dir) -> value            (`impl Hittable for Vec<..>`, :420-427)
 *   VK_EVAL_LIGHT_RANDOM   lights[index].random(p) with xi[0..2] -> out_d */
enum { VK_EVAL_BOUNCE = 0, VK_EVAL_BOUNCE_LEGACY = 1, VK_EVAL_TEXTURE = 2, VK_EVAL_LIGHTS_PDF = 3, VK_EVAL_LIGHT_RANDOM = 4 };
typedef struct vk_eval {
    uint32_t op;
    uint32_t index;                     /* BOUNCE*: material | TEXTURE: texture | LIGHT_RANDOM: entry of the light list */
    float ray_o[3], ray_d[3], ray_time; /* BOUNCE*: the ray that was traced                                              */
    float p[3], normal[3], t, u, v;     /* BOUNCE*: the HitRec | TEXTURE: u, v, p | LIGHT*: p = the origin               */
    uint32_t front;
    float dir[3];                       /* LIGHTS_PDF: the direction                                                    */
    uint32_t xi[5];                     /* BOUNCE*: xi[0..3] the bounce's four 32-bit variates (gen::<f32>() takes the
                                           top 24 bits, gen_range the top 23), xi[4] SpecDiffuse's choice;
                                           LIGHT_RANDOM: xi[0..2]                                                       */
    /* results */
    uint32_t alive;                     /* BOUNCE*: the path continues with (out_o, out_d, out_time)                     */
    uint32_t valid;                     /* BOUNCE*: 0 = the reference's sample is non-finite here (dropped, main.rs:192) */
    float out_o[3], out_d[3], out_time;
    float beta[3];                      /* BOUNCE*: path weight after the bounce | TEXTURE: the colour                   */
    float L[3];                         /* BOUNCE*: radiance added by this hit                                           */
    float value;                        /* LIGHTS_PDF                                                                   */
} vk_eval; /* 43 words */
int vk_eval_batch(vk_ctx* ctx, vk_eval* recs, size_t n, uint32_t flags /* VK_FLAG_STRICT_MATH */);

/* Microbenchmarks for the roofline denominators the driver does not measure:
 * dependent-free FFMA throughput (TFLOP/s) and L2-resident read bandwidth (GB/s). */
int vk_measure_peaks(vk_ctx* ctx, float* fp32_tflops, float* l2_gbs);

/* Device properties for reporting. */
int vk_device_info(vk_ctx* ctx, int* sm_count, int* clock_khz, char* name, size_t name_len);

#ifdef __cplusplus
}
#endif
#endif /* VECCHIO_GPU_H */
