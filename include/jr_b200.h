/*
 * jr_b200.h -- C ABI of the B200-native rasterisation path of jaxrenderer.
 *
 * Drop-in boundary: `renderer.pipeline.render(camera, shader, buffers,
 * face_indices, extra, loop_unroll)` (reference renderer/pipeline.py:470-537)
 * for the seven built-in shaders, and its reverse-mode derivative (the
 * reference obtains it from jax.grad, tests/smoke_test_grad.py:92-128).
 * A host binding (ctypes here; XLA-FFI handler for a JAX host, see
 * INTEGRATION.md) fills one JrRenderArgs and calls one entry point per render.
 *
 * Rules
 *  - every pointer is DEVICE memory owned by the caller; dense, row-major,
 *    fp32 / int32 (the reference's dtypes without x64);
 *  - every array argument is a JrF32/JrI32 {ptr, batch_stride}: element b of
 *    the batch starts at ptr + b * batch_stride ELEMENTS; batch_stride == 0
 *    broadcasts one array to the whole batch (jax.vmap in_axes=None);
 *  - buffers are x-major: zbuffer[b][x][y], canvas[b][x][y][c]
 *    (renderer/types.py:148-155), updated IN PLACE (the reference donates
 *    them, pipeline.py:466);
 *  - calls are asynchronous on `stream`, never allocate, never synchronise,
 *    never throw; they return 0 or a negative JrStatus;
 *  - re-entrant: no global state (a launch counter for benchmarks aside);
 *  - limits: W, H <= 32767; B <= 65535 for canvases larger than one shared-memory tile, for every
 *    non-depth shader (attribute-record stage) and for jr_render_backward (the batch is a grid
 *    dimension there); B * W * H < 2^30 for backward.  Everything a call needs (workspace size,
 *    limits) is validated before its first launch: a failing call enqueues nothing.
 */
#ifndef JR_B200_H_
#define JR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JR_ABI_VERSION 7

typedef void* jr_stream_t; /* cudaStream_t */

typedef enum JrShader {
  JR_DEPTH = 0,                   /* shaders/depth.py:38-61 */
  JR_GOURAUD = 1,                 /* shaders/gouraud.py:49-124 */
  JR_GOURAUD_TEXTURE = 2,         /* shaders/gouraud_texture.py:55-167 */
  JR_PHONG = 3,                   /* shaders/phong.py:65-187 */
  JR_PHONG_DARBOUX = 4,           /* shaders/phong_darboux.py:106-316 */
  JR_PHONG_REFLECTION = 5,        /* shaders/phong_reflection.py:91-258 */
  JR_PHONG_REFLECTION_SHADOW = 6, /* shaders/phong_reflection_shadow.py:99-295 */
  JR_NUM_SHADERS = 7
} JrShader;

typedef enum JrStatus {
  JR_OK = 0,
  JR_ERR_NULL = -1,        /* a required pointer is NULL */
  JR_ERR_DIMS = -2,        /* non-positive / inconsistent dimensions */
  JR_ERR_SHADER = -3,      /* unknown shader id */
  JR_ERR_WORKSPACE = -4,   /* workspace too small (see jr_workspace_bytes) */
  JR_ERR_UNSUPPORTED = -5, /* combination not supported */
  JR_ERR_CUDA = -6         /* a launch failed (cudaGetLastError) */
} JrStatus;

typedef struct JrF32 { const float* ptr; long long batch_stride; } JrF32;
typedef struct JrI32 { const int32_t* ptr; long long batch_stride; } JrI32;
typedef struct JrF32Out { float* ptr; long long batch_stride; } JrF32Out;

/*
 * One render call == reference `pipeline.render` vmapped over B images.
 * Fields a shader does not read may be left zero.
 */
typedef struct JrRenderArgs {
  int32_t shader;          /* JrShader */
  int32_t B, W, H;         /* batch, canvas width (axis 0), height (axis 1) */
  int32_t T;               /* triangles: face_indices (T,3) */
  int32_t n_pos;           /* rows of `position` (= extra[0].shape[0], pipeline.py:488-491) */
  int32_t n_nrm, n_uv;     /* rows of `normal`, `uv` */

  /* Camera (geometry.py:205-225): only the three matrices the path reads. */
  JrF32 world_to_clip;     /* (4,4) */
  JrF32 viewport;          /* (4,4) */
  JrF32 world_to_eye_norm; /* (4,4)  S4-S7 */

  /* Geometry.  `faces` indexes `position` (and `colour`).  faces_norm /
   * faces_uv / faces_tex index normal / uv / texture_index; NULL ptr means
   * "same as faces" (the generic pipeline.render call).  Renderer.render
   * (renderer.py:277-296) passes the model's three index buffers instead of
   * materialising the corner-expanded copies. */
  JrF32 position;          /* (n_pos,3) world space */
  JrI32 faces;             /* (T,3) */
  JrF32 normal;            /* (n_nrm,3) */
  JrI32 faces_norm;        /* (T,3) or NULL */
  JrF32 uv;                /* (n_uv,2) */
  JrI32 faces_uv;          /* (T,3) or NULL */
  JrF32 colour;            /* (n_pos,3)  S2 */

  /* Light (types.py:134-141; renderer.py:298-327). */
  JrF32 light_direction;   /* (3)  S2-S5 */
  JrF32 light_colour;      /* (3) */
  JrF32 light_dir_eye;     /* (3)  S6-S7 */
  JrF32 ambient, diffuse, specular; /* (3) each, S6-S7 */

  /* Maps. */
  JrF32 texture;           /* (tex_w, tex_h, 3) */
  int32_t tex_w, tex_h;
  JrF32 specular_map;      /* (spec_w, spec_h)  S6-S7 */
  int32_t spec_w, spec_h;
  JrF32 normal_map;        /* (tex_w, tex_h, 3)  S5 */
  JrI32 texture_shape;     /* (n_objects,2)  S6-S7 */
  int32_t n_objects;
  JrI32 texture_index;     /* (n_texidx)  S6-S7, read at faces_tex[t][0] */
  JrI32 faces_tex;         /* (T,3) or NULL */
  int32_t n_texidx;
  int32_t texture_offset;  /* model.py:552 */
  JrI32 id_to_face;        /* (n_pos)  S5 */
  JrI32 faces_indices;     /* (n_faces_indices,3)  S5 */
  int32_t n_faces_indices;

  /* Shadow (shadow.py:27-38)  S7. */
  JrF32 shadow_map;        /* (shadow_w, shadow_h) */
  int32_t shadow_w, shadow_h;
  JrF32 shadow_strength;   /* (3) */
  JrF32 shadow_world_to_clip; /* (4,4) light camera */
  JrF32 shadow_viewport;      /* (4,4) */

  /* Buffers, in/out (B,W,H) and (B,W,H,3).  canvas may be NULL for JR_DEPTH. */
  float* zbuffer;
  float* canvas;
  /* G-buffer out (B,W,H) int32: triangle written at each pixel, -1 if the
   * pixel kept its old value.  Required (it is what shading and backward
   * consume); may be NULL for JR_DEPTH when no gradient is wanted. */
  int32_t* tri_id;

  void* workspace;         /* >= jr_workspace_bytes(args) bytes, or NULL if that is 0 */
  size_t workspace_bytes;

  /* Instanced geometry: the world-space merge of `merge_objects` (renderer/model.py:447-555)
   * evaluated INSIDE the render kernels instead of being materialised per image.  Active when
   * inst_transform.ptr != NULL; then
   *   `position` holds the LOCAL-space vertices of all objects, concatenated (usually shared by the batch:
   *   batch_stride 0), and vertex i of image b is
   *        to_cartesian(to_homogeneous(position[i] * inst_scaling[o]) @ inst_transform[o]^T),  o = inst_vert_object[i]
   *   (model.py:489-499) -- the arithmetic of jr_merge_objects, so results equal the merged path bit for bit;
   *   `normal` holds the LOCAL normals and normal j is  ((normal[j] / f1) @ R^T) / f2  with R = inst_normal_matrix[o]
   *   (= inverse(transform)^T, host-computed), o = inst_norm_object[j] and (f1, f2) = inst_norm_scale[o]: the two
   *   whole-array (Frobenius) normalisations of model.py:517-530, produced by jr_instance_norm_scales.
   * Forward only (jr_render_backward wants materialised arrays); not with JR_PHONG_DARBOUX. */
  JrI32 inst_vert_object;    /* (n_pos) */
  JrF32 inst_scaling;        /* (n_inst,3) */
  JrF32 inst_transform;      /* (n_inst,4,4) */
  JrI32 inst_norm_object;    /* (n_nrm)  non-depth shaders */
  JrF32 inst_normal_matrix;  /* (n_inst,4,4) */
  JrF32 inst_norm_scale;     /* (n_inst,2) */
  int32_t n_inst;

  /* Display epilogue fused into the shading store (utils.py:79-98 + the uint8 cast of every published
   * benchmark, notebooks/32x32/A100.ipynb:326-337): when canvas_u8 != NULL the shading kernels (all shaders but
   * JR_DEPTH and JR_PHONG_DARBOUX) write  uint8(clamp(colour, 0, 1) * 255)  for EVERY pixel into canvas_u8 (B,H,W,3),
   * transposed and vertically flipped (row H-1-y, column x), instead of the fp32 canvas: pixels the render does not
   * write show the incoming `canvas` value if canvas != NULL, else canvas_u8_background.  z-buffer and tri_id are
   * written as usual; the fp32 canvas is then only read. */
  uint8_t* canvas_u8;
  float canvas_u8_background[3];

  /* Depth epilogue, JR_DEPTH only (the shadow-map pass, shadow.py:106-116, in ONE launch): every depth the
   * kernel writes is `z + depth_offset` (one rounded fp32 add; 0 = off), and when depth_fill != 0 the pixels no
   * triangle covers are written too, with `depth_fill_value + depth_offset` -- so the caller neither pre-fills the
   * map (renderer.py:349-354 fills it with the largest float) nor adds the offset afterwards. */
  float depth_offset;
  int32_t depth_fill;
  float depth_fill_value;

  /* Optional measurement counters (device, 8 x uint64, caller zero-fills; NULL = off, the normal case).
   * When set, the visibility kernels run their counting variant and ADD: [0] triangles visited,
   * [1] triangles passed by the filter phase, [2] triangles kept by the exact cull, [3] N_test = edge-function
   * evaluations (pixel x triangle tests; the reference's count is W*H*T, pipeline.py:163-279), [4] fragments
   * that passed the inside test, [5] exact-phase warp rounds, [6] batches. */
  unsigned long long* stats;
} JrRenderArgs;

/*
 * Cotangents.  Inputs: d_zbuffer (B,W,H), d_canvas (B,W,H,3) (either may be
 * NULL == zero).  Outputs are ACCUMULATED INTO (caller zero-fills) and may
 * be NULL when not wanted.  Batch strides follow the forward argument of the
 * same name: a broadcast input (stride 0) receives the batch-summed gradient.
 * d_zbuffer / d_canvas are overwritten in place with the cotangent w.r.t. the
 * INCOMING buffers (d_old = d_new * (1 - keep), pipeline.py:420-437).
 */
typedef struct JrGradArgs {
  float* d_zbuffer;            /* in: d/d z_out; out: d/d z_in   (B,W,H) */
  float* d_canvas;             /* in: d/d canvas_out; out: d/d canvas_in */
  JrF32Out d_position;         /* (n_pos,3) */
  JrF32Out d_normal;           /* (n_nrm,3) */
  JrF32Out d_colour;           /* (n_pos,3) */
  JrF32Out d_world_to_clip;    /* (4,4) */
  JrF32Out d_viewport;         /* (4,4) */
  JrF32Out d_world_to_eye_norm;/* (4,4) */
  JrF32Out d_light_direction, d_light_colour, d_light_dir_eye;
  JrF32Out d_ambient, d_diffuse, d_specular;
  JrF32Out d_texture;          /* (tex_w,tex_h,3) deterministic scatter */
  JrF32Out d_specular_map;     /* (spec_w,spec_h) */
  JrF32Out d_shadow_strength;  /* (3) */
  JrF32Out d_uv;               /* (n_uv,2)  phong_darboux only: through the tangent frame (phong_darboux.py:231-262) */
  JrF32Out d_normal_map;       /* (tex_w,tex_h,3) phong_darboux only, deterministic scatter */
  void* workspace;
  size_t workspace_bytes;
  int32_t no_buffer_grads;     /* non-zero: d_zbuffer / d_canvas are read-only inputs (may be NULL = zero); the
                                  cotangents of the incoming buffers are not wanted and not written */
} JrGradArgs;

int jr_abi_version(void);
const char* jr_strerror(int status);

/* Bytes of scratch `jr_render_forward` needs for these dimensions. */
size_t jr_workspace_bytes(const JrRenderArgs* args);
/* Bytes of scratch `jr_render_backward` needs. */
size_t jr_backward_workspace_bytes(const JrRenderArgs* args, const JrGradArgs* grads);

/* pipeline.render, forward (pipeline.py:470-537), any built-in shader. */
int jr_render_forward(const JrRenderArgs* args, jr_stream_t stream);
/* Reverse mode through fixed visibility (SURVEY 8a row BWD / Q9). */
int jr_render_backward(const JrRenderArgs* args, const JrGradArgs* grads, jr_stream_t stream);

/* Per-shader entry points (what an XLA-FFI target per shader binds to);
 * each checks args->shader and forwards to jr_render_forward/backward. */
int jr_depth_forward(const JrRenderArgs*, jr_stream_t);
int jr_gouraud_forward(const JrRenderArgs*, jr_stream_t);
int jr_gouraud_texture_forward(const JrRenderArgs*, jr_stream_t);
int jr_phong_forward(const JrRenderArgs*, jr_stream_t);
int jr_phong_darboux_forward(const JrRenderArgs*, jr_stream_t);
int jr_phong_reflection_forward(const JrRenderArgs*, jr_stream_t);
int jr_phong_reflection_shadow_forward(const JrRenderArgs*, jr_stream_t);

/* Shadow-map epilogue of Shadow.render_shadow_map (shadow.py:116):
 * shadow_map[i] += offset over n floats. */
int jr_add_scalar(float* data, long long n, float value, jr_stream_t stream);

/* Output epilogue: canvas (B,W,H,3) fp32 -> (B,H,W,3) uint8,
 * clamp(0,1)*255, transposed and vertically flipped (utils.py:79-98). */
int jr_canvas_to_uint8_display(const float* canvas, uint8_t* out, int B, int W, int H,
                               jr_stream_t stream);

/*
 * merge_objects (renderer/model.py:447-555), fused: world-space vertices and normals of all
 * objects of all batch elements in two launches instead of ~15 small framework ops per object
 * (first step).  Meshes are given concatenated in local space with an object id per
 * vertex / normal; `transform`, `normal_matrix` (= inverse(transform)^T, computed by the host) are
 * (n_objects,4,4) and `scaling` (n_objects,3), each optionally batched.
 *   verts  = to_cartesian(to_homogeneous(v * scaling) @ transform^T)            (model.py:489-499)
 *   norms  = apply_vec(n, normal_matrix) with the reference's whole-array (Frobenius)
 *            normalisation per object before and after the rotation             (model.py:517-530)
 */
typedef struct JrMergeArgs {
  int32_t B, n_objects, n_verts, n_norms;
  JrF32 local_verts;      /* (n_verts,3) */
  JrF32 local_norms;      /* (n_norms,3) */
  JrI32 vert_object;      /* (n_verts) object id of each vertex */
  JrI32 norm_start;       /* (n_objects+1) first normal of each object, prefix sums */
  JrF32 scaling;          /* (n_objects,3) */
  JrF32 transform;        /* (n_objects,4,4) */
  JrF32 normal_matrix;    /* (n_objects,4,4) */
  float* out_verts;       /* (B,n_verts,3) */
  float* out_norms;       /* (B,n_norms,3) */
} JrMergeArgs;
int jr_merge_objects(const JrMergeArgs* args, jr_stream_t stream);
/* The two Frobenius normalisation constants (f1, f2) per (image, object) of the normal transform above, WITHOUT
 * writing any normal: out_scales (B, n_objects, 2).  Same reductions, hence the same bits, as jr_merge_objects
 * (out_verts / out_norms of `args` are ignored).  Feeds JrRenderArgs.inst_norm_scale. */
int jr_instance_norm_scales(const JrMergeArgs* args, float* out_scales, jr_stream_t stream);

/* Fused camera construction: all 8 matrices of `Camera` (geometry.py:205-278) per
 * batch element in ONE launch.
 *   mode JR_CAMERA_PERSPECTIVE  Renderer.create_camera_from_parameters (renderer.py:141-196):
 *        lookAt view + analytic inverse (geometry.py:536-636), gluPerspective projection with
 *        aspect = tan(hfov/2) / tan(vfov/2) (geometry.py:638-684), viewport matrix (geometry.py:813-845).
 *        params row: position(3) target(3) up(3) vfov hfov near far viewWidth viewHeight viewDepth
 *   mode JR_CAMERA_LIGHT        the orthographic light camera of Shadow.render_shadow_map
 *        (shadow.py:73-98): eye = centre + light_direction * distance, glOrtho (geometry.py:720-763),
 *        viewport matrix taken from `viewport` (the main camera's).
 *        params row: centre(3) light_direction(3) up(3) distance left right bottom top near far
 * Both: world_to_clip = projection @ view, world_to_eye_norm = view_inv^T, world_to_screen =
 * (viewport @ projection) @ view, screen_to_world = (view_inv @ projection_inv) @ viewport_inv with the
 * closed-form inverses of geometry.py:472-511, :686-718 chosen by isclose(projection[3][3], 0)
 * (geometry.py:248-262).  out: (8, B, 4, 4) in the field order of `Camera`:
 * view, projection, viewport, world_to_clip, world_to_eye_norm, world_to_screen, view_inv, screen_to_world. */
#define JR_CAMERA_PERSPECTIVE 0
#define JR_CAMERA_LIGHT 1
typedef struct JrCameraArgs {
  int32_t B, mode;
  JrF32 params;     /* (16) per batch element, see above */
  JrF32 viewport;   /* (4,4), JR_CAMERA_LIGHT only */
  float* out;       /* (8,B,4,4) */
} JrCameraArgs;
int jr_camera_build(const JrCameraArgs* args, jr_stream_t stream);
/* Reverse mode of jr_camera_build (the reference differentiates its jnp builders, renderer.py:141-196,
 * shadow.py:84-103): d_out (8,B,4,4) = cotangents of the 8 matrices -> d_params (B,16) and, JR_CAMERA_LIGHT with
 * d_viewport != NULL, d_viewport (B,4,4).  `args->out` is ignored.  Exact derivative of the kernel's own formulas
 * (the same code evaluated on dual numbers), not a finite difference. */
int jr_camera_vjp(const JrCameraArgs* args, const float* d_out, float* d_params, float* d_viewport, jr_stream_t stream);

/* Test aid: audit of the conservative culls (the phase-A filter of the single-tile kernel and the bbox margin of
 * the exact phase / binned setup).  One warp per triangle brute-forces every pixel with the exact edge functions
 * (cost T*W*H per image: small scenes only) and ADDS to counters (device, 5 x uint64, caller zero-fills):
 * [0] triangles the filter drops although the exact cull keeps them, [1] pixels a kept triangle covers outside its
 * rasterised bbox, [2] pixels covered by triangles the bbox cull rejects, [3] triangles kept, [4] inside pixels.
 * [0], [1], [2] must be 0 (tests/test_gpu_fuzz.py, tests/test_gpu_full_size.py). */
int jr_debug_audit_cull(const JrRenderArgs* args, unsigned long long* counters, jr_stream_t stream);

/* Introspection for benchmarks: number of kernel launches issued by this
 * library since load (monotonic). */
long long jr_launch_count(void);

/* Measurement aid (ABI v7): per-kernel device times of the calls made through this library.  While switched on,
 * every entry point records a CUDA event on its stream at its start and after each of its launches;
 * jr_debug_kernel_times waits for those events, writes (kernel name, milliseconds since the event before it on the
 * same call) in launch order into out[0 .. min(n, capacity)), forgets them and returns n (or a negative JrStatus).
 * bench.py builds the per-kernel rooflines of its secondary lines from it.  Process-wide switch, at most
 * JR_KERNEL_TIMES_MAX marks are kept between two reads; do NOT leave it on while capturing a CUDA graph (event
 * records with timing cannot be captured) or while timing a step from outside (the events serialise nothing, but
 * each costs a host call).  Switching (either way) forgets what was recorded. */
#define JR_KERNEL_TIMES_MAX 4096
typedef struct JrKernelTime {
  char name[24];
  float ms;
  int32_t call; /* index of the entry-point call the launch belongs to (0, 1, ... since the switch-on / last read) */
} JrKernelTime;
int jr_debug_kernel_timing(int enable);
int jr_debug_kernel_times(JrKernelTime* out, int capacity);

#ifdef __cplusplus
}
#endif
#endif /* JR_B200_H_ */
