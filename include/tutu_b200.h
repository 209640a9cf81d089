/*
 * tutu_b200.h — C ABI of libtutu_b200.so, the B200-native path-tracing core for TutuRenderer.
 *
 * This is the drop-in boundary.  The reference has no FFI of its own: its two
 * plugin interfaces are C++ virtuals that are `new`-ed by name in Renderer::Renderer
 * (reference include/Renderer.hpp:35-54).  The entry points below are what C++ adapters deriving
 * from those interfaces bind (see include/tutu_adapters.hpp and INTEGRATION.md):
 *
 *   IIntersectStrategy::UpdateInter      (reference include/IIntersectStrategy.h:10-11,
 *                                         BVHStrategy.hpp:8-11 -> BVH.hpp:145 getIntersection)
 *        -> tutu_trace_closest
 *   isShadowRayBlocked / hasIntersection (reference include/IIntegrator.hpp:135-153, BVH.hpp:170-194)
 *        -> tutu_trace_any
 *   IIntegrator::integrate, PathTracing  (reference include/IIntegrator.hpp:17-24,
 *                                         PathTracing.hpp:352-475,485-516,136-279)
 *        -> tutu_render_path
 *   IIntegrator::integrate, BDPT         (reference include/BDPT.hpp:395-674,679-900,70-390)
 *        -> tutu_render_bdpt
 *   Scene / BVHAccel / PPMGenerator state read by the integrator
 *                                        (reference include/Scene.hpp:16-35, BVH.hpp:15-23,47-123,
 *                                         PPMGenerator.hpp:36-53,317-324, Camera.hpp:81-97)
 *        -> TutuSceneDesc + tutu_scene_upload (+ tutu_bvh_build for hosts without a reference tree)
 *
 * Plain pointers and sizes only; no C++/torch types.  All functions return TUTU_OK (0) or a
 * negative TUTU_E_* code; tutu_last_error() gives the message.  Nothing here ever falls back to
 * a CPU implementation: a missing GPU/driver is an error (TUTU_E_CUDA).
 */
#ifndef TUTU_B200_H
#define TUTU_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TUTU_ABI_VERSION 2

/* status codes */
#define TUTU_OK 0
#define TUTU_E_INVALID (-1) /* bad argument / malformed scene */
#define TUTU_E_CUDA (-2)    /* CUDA runtime or driver error (includes "no device") */
#define TUTU_E_STATE (-3)   /* call order (e.g. trace before scene upload) */
#define TUTU_E_IO (-4)      /* scene file IO */
#define TUTU_E_NOMEM (-5)

/* primitive kinds — reference include/Object.hpp:9 (OBJTYPE {TRIANGLE, SPEHRE}) */
#define TUTU_PRIM_TRIANGLE 0
#define TUTU_PRIM_SPHERE 1

/* material kinds, same order as reference include/Material.hpp:9-16 (MaterialType) */
#define TUTU_MAT_LAMBERTIAN 0
#define TUTU_MAT_PERFECT_REFLECTIVE 1
#define TUTU_MAT_PERFECT_REFRACTIVE 2
#define TUTU_MAT_MICROFACET_R 3
#define TUTU_MAT_MICROFACET_T 4
#define TUTU_MAT_UNLIT 5

/* texture channels — reference include/PPMGenerator.hpp:36-39 */
#define TUTU_TEX_DIFFUSE 0
#define TUTU_TEX_NORMAL 1
#define TUTU_TEX_ROUGHNESS 2
#define TUTU_TEX_METALLIC 3

/* reference include/Material.hpp:21-30 (same field order; 56 bytes) */
typedef struct TutuMaterial {
  float diffuse[3];
  float specular[3];
  float emission[3];
  int32_t type;
  float alpha;
  float eta;
  float roughness;
  float metallic;
} TutuMaterial;

/* One reference Object (Triangle.hpp:11-18 / Sphere.hpp:8-9 + Object.hpp:26-35), 124 bytes.
 * Triangle: v = v0,v1,v2; n = n0,n1,n2; uv = uv0,uv1,uv2.
 * Sphere:   v[0..2] = centre, v[3] = radius; n, uv ignored. */
typedef struct TutuPrim {
  int32_t type;
  float v[9];
  float n[9];
  float uv[6];
  int32_t material;      /* index into TutuSceneDesc.materials */
  int32_t tex_active;    /* Object::isTextureActivated */
  int32_t tex_diffuse;   /* Object::textureIndex,      -1 = none */
  int32_t tex_normal;    /* Object::normalMapIndex,    -1 = none */
  int32_t tex_roughness; /* Object::roughnessMapIndex, -1 = none */
  int32_t tex_metallic;  /* Object::metallicMapIndex,  -1 = none */
} TutuPrim;

/* Pre-order export of the reference's BVHNode tree (BVH.hpp:15-23): node 0 is the root, an
 * internal node has left,right >= 0 and prim = -1; a leaf has left = right = -1 and
 * prim = index into TutuSceneDesc.prims.  Bounds are recomputed by the library with the
 * reference's own Union/fmin/fmax (BoundBox.hpp:97-124), which is exact. */
typedef struct TutuBvhNode {
  int32_t left;
  int32_t right;
  int32_t prim;
} TutuBvhNode;

/* reference include/Texture.hpp:9-14: row-major rgb floats, 3 per texel */
typedef struct TutuTexture {
  int32_t width;
  int32_t height;
  const float* rgb;
} TutuTexture;

/* Raw camera inputs as the config file gives them (PPMGenerator.hpp:41-47, Camera.hpp:81-91);
 * the library derives the ray-generation constants exactly as PathTracing.hpp:357-391 does. */
typedef struct TutuCamera {
  float eye[3];
  float viewdir[3];
  float updir[3];
  int32_t hfov_deg;
  int32_t width;
  int32_t height;
  int32_t parallel_projection;
} TutuCamera;

typedef struct TutuSceneDesc {
  uint32_t struct_size; /* = sizeof(TutuSceneDesc) */
  uint32_t n_prims;
  const TutuPrim* prims; /* Scene::objList order */
  uint32_t n_materials;
  uint32_t n_bvh_nodes; /* 0 => the library builds the midpoint BVH itself (tutu_bvh_build) */
  const TutuMaterial* materials;
  const TutuBvhNode* bvh_nodes;
  const TutuTexture* tex[4]; /* indexed by TUTU_TEX_* */
  uint32_t n_tex[4];
  TutuCamera camera;
  float bkgcolor[3];
  float eta; /* scene index of refraction, PPMGenerator.hpp:48 */
} TutuSceneDesc;

/* Closest-hit record, 16 bytes.  prim = Scene::objList index, -1 on a miss (then t = FLT_MAX,
 * u = v = 0).  u,v are the barycentric weights of v1 and v2 (Triangle.hpp:47-49); 0 for spheres. */
typedef struct TutuHit {
  int32_t prim;
  float t;
  float u;
  float v;
} TutuHit;

/* Ray record, 32 bytes: {o.x,o.y,o.z,unused, d.x,d.y,d.z,tmax}.  tmax is only read by
 * tutu_trace_any* (the `dis` argument of hasIntersection, BVH.hpp:170). */
#define TUTU_RAY_FLOATS 8

typedef struct TutuSceneInfo {
  uint32_t n_prims;
  uint32_t n_nodes;      /* reference-tree node count (2*n_prims-1) */
  uint32_t n_inner;      /* device inner nodes (n_prims-1) */
  uint32_t depth;        /* reference-tree depth, root = 0 */
  uint32_t n_lights;     /* emissive prims, PPMGenerator.hpp:317-324 */
  uint32_t n_materials;
  uint32_t width, height;
  uint64_t device_bytes; /* HBM held by the uploaded scene */
  /* the tree the production walk descends for regular rays (not the reference's topology) */
  uint32_t trav_nodes;      /* inner nodes */
  uint32_t trav_depth;      /* deepest inner node, root = 1 */
  uint32_t trav_width;      /* children per node: 2 (binary SAH tree) or 8 (compressed wide tree) */
  uint32_t trav_node_bytes; /* bytes fetched per inner-node visit */
  uint32_t trav_leaf_bytes; /* bytes fetched per primitive test */
  uint32_t reserved;
} TutuSceneInfo;

/* Counters of the last tutu_render_path* call (diagnostics and the roofline arithmetic). */
typedef struct TutuRenderStats {
  uint64_t paths;        /* camera samples started */
  uint64_t extend_rays;  /* closest-hit rays traced */
  uint64_t shadow_rays;  /* any-hit rays traced */
  uint64_t shade_calls;  /* shading-vertex evaluations (BDPT: sub-path vertices stored) */
  uint64_t nan_samples;  /* samples dropped by the NaN filter (PathTracing.hpp:510) */
  uint64_t kernel_launches;
  uint64_t iterations;   /* wavefront iterations */
  float gpu_ms;          /* device time of the render, CUDA events on the render stream */
  float extend_ms, shade_ms, shadow_ms, other_ms; /* only filled when profiling is enabled */
} TutuRenderStats;

typedef struct TutuCtx TutuCtx;

/* ---- context ---------------------------------------------------------------------------- */
int tutu_abi_version(void);
/* device = CUDA ordinal.  Fails with TUTU_E_CUDA when no usable GPU exists (no CPU fallback). */
int tutu_ctx_create(int device, TutuCtx** out);
void tutu_ctx_destroy(TutuCtx* ctx);
/* Message of the last failing call on this ctx (or, with ctx == NULL, on this thread). */
const char* tutu_last_error(const TutuCtx* ctx);

/* ---- scene ------------------------------------------------------------------------------ */
/* Flatten + upload.  Host pointers are not retained past the call. */
int tutu_scene_upload(TutuCtx* ctx, const TutuSceneDesc* desc);
int tutu_scene_info(const TutuCtx* ctx, TutuSceneInfo* out);
/* Who builds the traversal tree for regular rays (a binary tree over the reference's leaves; the reference's own
 * topology is always kept for irregular rays and the literal walk).  TUTU_BUILD_HOST_SAH = binned surface-area
 * heuristic on the host (0.16 s for 10^6 primitives on 16 threads; 34 node visits per ray on configs[1]);
 * TUTU_BUILD_DEVICE_SAH = the same split rule on the GPU, level-synchronous (the same tree in 9.6 ms);
 * TUTU_BUILD_DEVICE_LBVH = linear BVH on the GPU (Morton keys, radix sort, Karras hierarchy, refit: 6.7 ms, but 91
 * node visits per ray); TUTU_BUILD_DEVICE_PLOC = the same Morton order merged bottom-up by surface area (parallel
 * locally-ordered clustering: 8.4 ms, 69 node visits per ray); TUTU_BUILD_AUTO = the device SAH builder for scenes
 * of more than 64 primitives, the host's below (DESIGN.md 5.8 has the measurements).  Takes effect at the next
 * tutu_scene_upload.  Hits do not depend on the choice (any tree with exact union boxes over the same leaves gives
 * the same answer); device-built trees deeper than 30 levels fall back to the host builder. */
#define TUTU_BUILD_AUTO 0
#define TUTU_BUILD_HOST_SAH 1
#define TUTU_BUILD_DEVICE_LBVH 2
#define TUTU_BUILD_DEVICE_PLOC 3 /* GPU: Morton sort + parallel locally-ordered clustering (bottom-up merges by surface area) */
#define TUTU_BUILD_DEVICE_SAH 4  /* GPU: the host builder's binned-SAH split rule, level-synchronous (the same tree) */
int tutu_scene_builder(TutuCtx* ctx, int builder);
/* Wall-clock breakdown of the last tutu_scene_upload (milliseconds). */
typedef struct TutuUploadStats {
  float total_ms;
  float flatten_ms;    /* host: validation, DFS slots, leaf records, reference-topology nodes (+ midpoint build if no tree was given) */
  float tree_build_ms; /* traversal tree for regular rays: host SAH build, or device LBVH (incl. its two small uploads) */
  float h2d_ms;        /* copies of the flattened arrays */
  int32_t builder;     /* TUTU_BUILD_HOST_SAH or one of TUTU_BUILD_DEVICE_*: what was actually used */
  uint32_t tree_depth;
} TutuUploadStats;
int tutu_upload_stats(const TutuCtx* ctx, TutuUploadStats* out);
/* Change the frame size / camera without re-uploading geometry. */
int tutu_scene_set_camera(TutuCtx* ctx, const TutuCamera* cam);

/* ---- ray batches: IIntersectStrategy::UpdateInter / hasIntersection --------------------- */
/* Host buffers: H2D copy of rays, traversal, D2H copy of results, synchronous. */
int tutu_trace_closest(TutuCtx* ctx, const float* rays, uint64_t n_rays, TutuHit* hits_out);
int tutu_trace_any(TutuCtx* ctx, const float* rays, uint64_t n_rays, uint8_t* blocked_out);
/* Device buffers (CUDA device pointers valid on the ctx's device), asynchronous on `stream`
 * (a cudaStream_t / CUstream passed as void*, NULL = the ctx's own stream).  Batches may be enqueued on
 * different streams: the context's work cursor and binning scratch are handed from one batch to the next with
 * an event, so batches of one context run one after the other on the device. */
int tutu_trace_closest_device(TutuCtx* ctx, const float* d_rays, uint64_t n_rays,
                              TutuHit* d_hits_out, void* stream);
int tutu_trace_any_device(TutuCtx* ctx, const float* d_rays, uint64_t n_rays,
                          uint8_t* d_blocked_out, void* stream);
/* Traversal variant: 0 = ordered + t-pruned walk, large batches traced in a coherent order
 * (counting sort by entry cell + direction bin; results are order independent) — the default;
 * 3 = the same walk in the caller's ray order; 1 = unpruned both-children walk that mirrors the
 * reference's recursion literally (used as a second opinion by the tests); 4 = the general tree walk also on scenes
 * of <= 32 primitives, for ray batches and renders (otherwise their flat leaf-box kernels are used);
 * 6 = as 0, but regular rays (every 1/d finite) walk the compressed 8-wide collapse of the traversal tree (8-bit
 * child boxes quantised outwards, decoded exactly; built on the first use).  Same hits bit for bit; on a B200 it
 * is slower than the binary walk, so it is an option. */
int tutu_set_traversal_mode(TutuCtx* ctx, int mode);
/* Where the tree kernels keep their per-ray traversal stack: 0 = decided per scene (the default: local memory while the
 * traversal arrays — both node arrays and the leaf geometry — are at most 32 MB and live in L1/L2 next to it, shared
 * memory above that, where the stack's local-memory lines would compete with the node fetches),
 * 1 = shared memory, 2 = local memory.  Same results either way; the tests run both. */
int tutu_traversal_stack(TutuCtx* ctx, int where);
/* Queue tracers of tutu_render_bdpt on scenes with a tree (closest hit for the sub-path walks, any hit for the connections):
 * 1 = every warp walks one packet of 32 queue entries at a time; 2 = persistent lanes: a warp refills its idle lanes from the
 * queue and runs its lanes' node steps and primitive tests as separate phases; 0 (the default) = measured
 * per scene: the first render of at least 8 batches runs one batch with each, alone, and the faster one is kept until the
 * next tutu_scene_upload.  Same results either way.  tutu_bdpt_queue_tracer_measured reports that measurement: *tracer =
 * -1 (none yet), 0 (packets) or 1 (lanes), and the two batch times in ms. */
int tutu_bdpt_queue_tracer(TutuCtx* ctx, int tracer);
int tutu_bdpt_queue_tracer_measured(const TutuCtx* ctx, int* tracer, float* ms_packets, float* ms_lanes);
/* Visit counters for the algorithmic-bytes figure: traces the batch with counting kernels and
 * returns total inner-node fetches and primitive tests. */
int tutu_trace_count_visits(TutuCtx* ctx, const float* d_rays, uint64_t n_rays, int any_hit,
                            uint64_t* nodes_out, uint64_t* prims_out);

/* ---- PathTracing::integrate ------------------------------------------------------------- */
/* Whole render through host memory: spp samples per pixel, linear radiance written to
 * rgb_out[(y*width+x)*3 + c] exactly where the reference writes cam.FrameBuffer.rgb
 * (PathTracing.hpp:501,513).  Synchronous. */
int tutu_render_path(TutuCtx* ctx, uint32_t spp, uint64_t seed, float* rgb_out);
/* Multi-GPU building block: accumulate samples [sample_begin, sample_begin+sample_count) of
 * every pixel as SUMS (NaN samples dropped) into d_accum (width*height*3 floats on the device,
 * not cleared).  The work is ordered after everything already enqueued on `stream`, and `stream`
 * is ordered after it; the call itself returns once the samples are done (it polls the wavefront).  The caller reduces d_accum across ranks and then calls
 * tutu_finalize_device with inv_spp = 1/total_spp. */
int tutu_render_path_accumulate_device(TutuCtx* ctx, uint32_t sample_begin, uint32_t sample_count,
                                       uint64_t seed, float* d_accum, void* stream);
int tutu_finalize_device(TutuCtx* ctx, const float* d_accum, float inv_spp, float* d_rgb_out,
                         void* stream);
/* ---- BDPT::integrate (reference include/BDPT.hpp:395-674 -> sub_render_bdpt :679-900) -------- */
/* Whole bidirectional render through host memory.  rgb_out = bkgcolor + sum of the weighted
 * (s,t) strategies / spp, exactly where the reference adds them (cam.FrameBuffer.rgb, pre-filled
 * with bkgcolor by Camera::initialize, Camera.hpp:28; BDPT.hpp:823,891).  Synchronous. */
int tutu_render_bdpt(TutuCtx* ctx, uint32_t spp, uint64_t seed, float* rgb_out);
/* Multi-GPU building blocks, as for the path tracer: accumulate the SUMS of samples
 * [sample_begin, sample_begin+sample_count) into d_accum (not cleared; bkgcolor not included), then
 * after the cross-rank reduce tutu_finalize_bdpt_device writes bkgcolor + d_accum * inv_spp. */
int tutu_render_bdpt_accumulate_device(TutuCtx* ctx, uint32_t sample_begin, uint32_t sample_count,
                                       uint64_t seed, float* d_accum, void* stream);
int tutu_finalize_bdpt_device(TutuCtx* ctx, const float* d_accum, float inv_spp, float* d_rgb_out,
                              void* stream);
int tutu_render_stats(const TutuCtx* ctx, TutuRenderStats* out);
/* Knobs: paths in flight per wavefront lane (0 = default: 32 Mi for scenes shaded in queue order — a single
 * untextured Lambertian material class, e.g. the Cornell box — else 16 Mi; BDPT: samples per batch, default
 * 4 Mi), number of interleaved wavefront lanes (0 = default: 1 resp. 2 for the same two cases), per-stage
 * event timing on/off. */
int tutu_render_configure(TutuCtx* ctx, uint64_t paths_in_flight, int lanes, int profile_stages);
/* Path-tracing pipeline: 0 = automatic (default), 1 = wavefront (queues in HBM, any scene), 2 =
 * register-resident persistent kernel (scenes of <= 32 primitives, whose geometry fits the kernel's
 * constant bank; a render on a larger scene then fails with TUTU_E_STATE).  Automatic takes 2 for renders
 * of fewer than 768 Ki paths on such scenes (one launch, no queue pools: 3.4x faster at 64x64 @ 16 spp) and
 * the wavefront otherwise (1.5x faster once its queues fill).  Both run the same vertex code
 * on the same random numbers: images agree to float noise (frame-buffer summation order, FMA contraction
 * per translation unit). */
int tutu_render_pipeline(TutuCtx* ctx, int pipeline);

/* ---- output stage: PPMGenerator::writePixel (reference include/PPMGenerator.hpp:812-845) ---- */
/* 8-bit quantisation of a linear radiance image exactly as the reference writes its PPM:
 * out = (int)(255 * pow(clamp(0, 1, c), gamma)) per channel, gamma = GAMMA_VAL = 0.78 (global.hpp:30);
 * gamma <= 0 selects the reference's non-gamma branch (255 * clamp).  NaN clamps to 1 like
 * std::max(lo, std::min(hi, v)) does (global.hpp:52-55).  rgb: width*height*3 floats, out: width*height*3
 * bytes.  Host buffers (H2D, kernel, D2H; synchronous) and device buffers (asynchronous). */
int tutu_quantize(TutuCtx* ctx, const float* rgb, uint64_t n_pixels, float gamma, uint8_t* out);
int tutu_quantize_device(TutuCtx* ctx, const float* d_rgb, uint64_t n_pixels, float gamma, uint8_t* d_out,
                         void* stream);

/* ---- output stage: Postprocessor (reference include/Postprocessor.hpp:29-197) ----------------- */
/* The reference's bloom / exposure post-process on a linear radiance image (its call is commented out in
 * every driver, src/main_cornellBox.cpp:82-84; the class is complete).  Parameters = the reference's
 * #defines (Postprocessor.hpp:10-14) and the literal in getEmmisiveTexture (:141). */
typedef struct TutuPostParams {
  float emissive_norm;    /* pixels with |rgb| > 3.f count as emissive (Postprocessor.hpp:141) */
  float strength;         /* STRENGTH 2: the brightest channel of an emissive pixel is rescaled to this */
  int32_t gaussian_loops; /* GAUSSIANLOOP 1: the extract is blurred 1 + gaussian_loops times */
  int32_t kernel_size;    /* KERNELSIZE 10: taps start at (int)(-kernel_size * 0.5), so 10 taps are -5..4 */
  float stddev;           /* STDDEV 30 */
  float exposure;         /* EXPOSURE 1.5 */
} TutuPostParams;
void tutu_post_params_default(TutuPostParams* out);
#define TUTU_POST_EXTRACT 1   /* getEmmisiveTexture */
#define TUTU_POST_BLUR 2      /* getGaussianBlurTexture, once (vertical pass, then horizontal) */
#define TUTU_POST_BLOOM 3     /* extract, 1 + loops blurs, add to the source (performPostProcess under BLOOM_ONLY) */
#define TUTU_POST_HDR 4       /* getHDRtexture: 1 - exp(-c * exposure) (performPostProcess under HDR_ONLY) */
#define TUTU_POST_HDR_BLOOM 5 /* bloom, then the tone map (performPostProcess under HDR_BLOOM) */
/* rgb, rgb_out: width*height*3 floats (rgb_out may equal rgb for the host entry point).  params = NULL takes
 * the reference's constants.  Host buffers (H2D, kernels, D2H; synchronous) and device buffers (asynchronous on
 * `stream`; d_rgb_out must not alias d_rgb).  Every read goes through Texture::getRGBat's index arithmetic
 * (Texture.hpp:18-39) exactly as in the reference, including its u == 0 wrap-around. */
int tutu_postprocess(TutuCtx* ctx, const float* rgb, uint32_t width, uint32_t height, int mode,
                     const TutuPostParams* params, float* rgb_out);
int tutu_postprocess_device(TutuCtx* ctx, const float* d_rgb, uint32_t width, uint32_t height, int mode,
                            const TutuPostParams* params, float* d_rgb_out, void* stream);

/* ---- host-side helpers (no GPU needed) -------------------------------------------------- */
/* PPM file from 8-bit pixels: binary = 0 writes the reference's ASCII P3 byte for byte
 * (PPMGenerator.hpp:804-809, 840-842: "P3\nW\nH\n255\n" then "r g b\n" per pixel), binary = 1
 * writes P6 (the author's TODO, README.md:49). */
int tutu_write_ppm(const char* path, uint32_t width, uint32_t height, const uint8_t* rgb8, int binary);
int tutu_write_png(const char* path, uint32_t width, uint32_t height, const uint8_t* rgb8); /* 8-bit RGB PNG */
/* Texture files for TutuTexture.rgb (scene authoring).  ASCII P3 is read exactly as the
 * reference's PPMGenerator::loadTexture does (PPMGenerator.hpp:1027-1084: texel = (r/max, g/max, b/max) in
 * fp32, rows top to bottom); binary P6 (maxval <= 65535) and PNG (grey, grey+alpha, RGB, RGBA at 8/16 bits,
 * grey at 1/2/4 bits, palette; alpha ignored; no Adam7) are what it cannot read (:1050).  normal_map != 0
 * applies the reference's `bump` recovery c * 2 - 1 (PPMGenerator.hpp:714-720).  *rgb_out holds
 * width*height*3 floats and is released with tutu_texture_free. */
int tutu_texture_load(const char* path, int normal_map, float** rgb_out, int32_t* width, int32_t* height);
void tutu_texture_free(float* rgb);
/* Midpoint BVH with the reference's split rule (BVH.hpp:47-123).  nodes_out must hold
 * 2*n_prims-1 entries (1 if n_prims <= 1). */
int tutu_bvh_build(const TutuPrim* prims, uint32_t n_prims, TutuBvhNode* nodes_out,
                   uint32_t* n_nodes_out);
/* Host-only self check of the traversal trees tutu_scene_upload would build for `desc` (the binned-SAH binary
 * tree over the reference's leaves and its compressed 8-wide collapse): every child box of every wide node,
 * decoded with the device's own arithmetic, must contain the exact boxes of all leaves below it, and every
 * leaf must be referenced exactly once.  violations == 0 is what the bit-exact hit parity of regular rays
 * rests on. */
typedef struct TutuTreeCheck {
  uint32_t n_leaves;
  uint32_t binary_nodes, binary_depth;
  uint32_t wide_nodes, wide_depth; /* 0 = no wide tree (degenerate scene: the device walks the binary tree) */
  uint64_t wide_children;          /* occupied slots over all wide nodes */
  uint64_t violations;
} TutuTreeCheck;
int tutu_traversal_tree_check(const TutuSceneDesc* desc, TutuTreeCheck* out);
/* Scene files (tests, benches, the oracle harness): a flat little-endian dump of TutuSceneDesc. */
typedef struct TutuSceneFile TutuSceneFile;
int tutu_scene_file_load(const char* path, TutuSceneFile** out);
const TutuSceneDesc* tutu_scene_file_desc(const TutuSceneFile* f);
void tutu_scene_file_free(TutuSceneFile* f);
int tutu_scene_file_save(const TutuSceneDesc* desc, const char* path);
/* Synthetic workload of BASELINE.json configs[1]: a G x G height-field (2*G*G triangles) and
 * ray batches against it.  prims_out must hold 2*G*G entries; rays_out n_rays*8 floats.
 * kind: 0 = top-down rays (coherent, ~all hit), 1 = incoherent rays from inside the bounds. */
int tutu_synth_heightfield(uint32_t G, uint64_t seed, TutuPrim* prims_out);
int tutu_synth_rays(int kind, uint64_t seed, uint64_t first, uint64_t n_rays, float* rays_out);
/* Debug aid (compute-sanitizer stand-in, tests/test_gpu_guards.py): in a library compiled with -DTUTU_GUARDS every
 * device allocation sits between two 64 KB bands of a byte pattern; this synchronises the device and counts the live
 * allocations and the band bytes that were overwritten (bands of allocations freed since start-up included).  The
 * shipped library has no bands and returns TUTU_E_STATE. */
int tutu_debug_guard_check(uint64_t* n_buffers, uint64_t* n_bad_bytes);
/* Self-test of that detector: zeroes n_bytes of the upper band of the oldest live allocation (guard builds only). */
int tutu_debug_guard_poke(uint32_t n_bytes);

#ifdef __cplusplus
}
#endif
#endif /* TUTU_B200_H */
