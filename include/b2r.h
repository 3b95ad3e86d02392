/* b2r.h -- C ABI of the B200-native frame pipeline for py-numpy-renderer.
 *
 * The reference (Denizantip/py-numpy-renderer) has no FFI of its own: its hot path is the body of
 * `Scene.render()` (obj/core.py:587-640) calling `rasterize` (obj/triangular.py:29-132), `general_shading`
 * (triangular.py:135-171), `shadow_volumes` (triangular.py:294-302), `resterize_quadrangle`
 * (triangular.py:319-368), `clipping` (obj/plane_intersection.py:59-86) and `fill_frame_from_skybox`
 * (obj/cube_map.py:83-101) on caller-owned NumPy arrays.  This header is the boundary a maintainer would bind
 * (ctypes stub in INTEGRATION.md) to replace exactly that body:
 *
 *   b2r_scene_create   <- the data `Scene.add_model(Model)` accumulates            (core.py:231-256, 584-585)
 *                         + `TextureMaps.register` / `parse_mtl` textures          (core.py:90-105, 321-348)
 *                         + `CubeMap(...)`                                         (cube_map.py:22-61)
 *   b2r_render         <- `Scene.render()`                                         (core.py:587-640)
 *   b2r_scene_reset_silhouette <- the persistent `model.silhouette` set            (core.py:251, 605)
 *
 * Plain pointers and sizes only; no torch / Python types.  All matrices are float64, row-major, ROW-VECTOR
 * convention (v' = v @ M) and are computed by the host with the reference's own NumPy expressions
 * (core.py:394-429), so the library never re-derives a camera.
 *
 * Error model: every call returns 0 on success, non-zero on failure; `b2r_last_error()` gives the message
 * (thread-local).  B2R_ERR_INDEX (2) from b2r_render / b2r_wait / b2r_sync means a texture lookup fell outside its map
 * (UV < -1), where the reference raises IndexError (core.py:138-143, 162-173).  The library never frees or keeps
 * caller memory: create() copies what it needs to the device.
 * There is NO CPU fallback: without a CUDA device every entry point except b2r_last_error/b2r_abi_version fails.
 *
 * State and threading: there is no process-global render state.  b2r_init(device) creates (or selects) the CONTEXT of
 * that CUDA device -- its streams, events and pinned staging -- and makes it the calling thread's current context; a
 * scene belongs to the context it was created on, so scenes on different devices coexist in one process and may be
 * driven from different threads.  Every entry point locks the context it acts on (entry points taking a scene: the
 * scene's; the others: the calling thread's current context, i.e. the device it last passed to b2r_init, else the most
 * recently initialised one).
 */
#ifndef B2R_H_
#define B2R_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2R_ABI_VERSION 6
#define B2R_MAX_POLY 12 /* a quad clipped by 6 planes has at most 10 vertices */

/* Light kinds: obj/lightning.py:4-7 */
enum { B2R_LIGHT_DIRECTIONAL = 0, B2R_LIGHT_POINT = 1, B2R_LIGHT_SPOT = 2 };
/* Texel decode: core.py:96-104 -- f32(u8/255) or f32(u8/255*2-1) */
enum { B2R_TEX_UNORM = 0, B2R_TEX_SNORM = 1 };
/* vertex storage of Model.vertices: float32 straight from the loader, float64 after an `@` chain (core.py:350-352) */
enum { B2R_F32 = 0, B2R_F64 = 1 };
/* Shading function applied at the call site of rasterize() (triangular.py:120-130).  The reference ships the call to
 * general_shading and, commented out beside it, flat_shading / gouraud / pbr (triangular.py:174-263): selecting one of
 * those reproduces what the reference renders with that line swapped in (both passes then write the same value, so
 * the shadow stencil has no visible effect). */
enum { B2R_SHADE_GENERAL = 0, B2R_SHADE_FLAT = 1, B2R_SHADE_GOURAUD = 2, B2R_SHADE_PBR = 3 };
/* background: core.py:595-600 */
enum { B2R_BG_COLOR = 0, B2R_BG_CUBEMAP = 1 };
/* per-face status of pass 3, triangular.py:15-20 (bit values of the reference's `Errors` Flag; 0 = rendered) */
enum { B2R_FACE_RENDERED = 0, B2R_FACE_BACK_FACE_CULLING = 1, B2R_FACE_WRONG_MIN_MAX = 2, B2R_FACE_EMPTY_B = 4,
       B2R_FACE_EMPTY_Z = 8, B2R_FACE_CLIPPED = 16 };

typedef struct b2r_texture_desc {
    const uint8_t* rgb; /* host, height*width*3, row-major, as PIL `convert('RGB')` yields (core.py:100-105) */
    int32_t height, width;
    int32_t decode;  /* B2R_TEX_UNORM | B2R_TEX_SNORM */
    int32_t tangent; /* normal maps: dtype metadata 'tangent' (core.py:94, 180) */
} b2r_texture_desc;

/* obj/materials.py:47-55 + the maps Face.get_object_color/get_specular/get_normals look up (core.py:145-189) */
typedef struct b2r_material {
    double Kd[3];
    double Ks[3];
    double Ns;
    int32_t map_Kd; /* index into the scene texture table or -1 */
    int32_t map_Ks;
    int32_t norm;
    int32_t reserved;
    double Pm, Pr;  /* metalness, roughness (materials.py:47-48), read by pbr() only */
    double Ka[3];   /* ambient colour (materials.py:49), the `ao` of pbr() */
} b2r_material;

/* One `Model` (core.py:231-256).  faces is the reference's (F,3,4) int32 array: per corner [v, vt, vn, mtl]. */
typedef struct b2r_model_desc {
    const void* vertices; /* (V,4) vertex_dtype */
    const void* uv;       /* (T,3) uv_dtype or NULL */
    const void* normals;  /* (N,3) normal_dtype or NULL */
    const int32_t* faces; /* (F,3,4) */
    const b2r_material* materials; /* indexed by faces[f][0][3] (out of range -> materials[0] = 'default') */
    int32_t n_vertices, n_uv, n_normals, n_faces, n_materials;
    int32_t vertex_dtype, uv_dtype, normal_dtype; /* B2R_F32 | B2R_F64 */
    int32_t clip;       /* Model.clip (triangular.py:80) */
    int32_t depth_test; /* Model.depth_test (triangular.py:117): 0 = the model's faces are z-TESTED but never write z.
                           Still order independent: zbuf = min/max over the writing faces only, and the face that colours
                           a pixel is the greatest index among {writing faces with z == zbuf} and {non-writing faces
                           whose z passes the test against the final zbuf} */
} b2r_model_desc;

/* CubeMap.textures after the constructor's flips/rotations (cube_map.py:22-44): 6 faces ordered
 * right,left,top,bottom,front,back, each size*size*3 uint8. */
typedef struct b2r_cubemap_desc {
    const uint8_t* faces;
    int32_t size;
    int32_t reserved;
} b2r_cubemap_desc;

/* Per camera view: everything `rasterize`/`resterize_quadrangle`/`fill_frame_from_skybox` read from the camera. */
typedef struct b2r_view {
    double mvp[16];       /* camera.MVP             core.py:419-421 */
    double mvp_dbg[16];   /* debug_camera.MVP       triangular.py:39 */
    double viewport[16];  /* camera.viewport        core.py:427-429, transformation.py:123-136 */
    double planes[24];    /* camera.frustum_planes  plane_intersection.py:43-56 */
    double sky_inv[16];   /* inv(lookat_without_translation @ projection)  cube_map.py:94-97 */
    double cam_pos[3];    /* camera.position        triangular.py:156 */
    double near_, far_;   /* camera.near/far        core.py:226-228 */
    int32_t system;           /* +1 RH, -1 LH      constants.py:29-31 */
    int32_t backface_culling; /* triangular.py:47 */
} b2r_view;

/* core.py:444-524 */
typedef struct b2r_light {
    double position[3];
    double direction[3]; /* normalize(position - center), core.py:364-366 */
    double color[3];
    double ambient[3];   /* ambient_strength * color, core.py:464 */
    double specular_strength, constant, linear, quadratic;
    double spot_cos_outer, spot_cos_inner; /* cos(20 deg), cos(10 deg): triangular.py:158-159 */
    int32_t type;
    int32_t reserved;
} b2r_light;

typedef struct b2r_frame_params {
    b2r_light light;
    float background[3]; /* used when bg_mode == B2R_BG_COLOR (core.py:597-600, already float32) */
    int32_t bg_mode;
    int32_t height, width; /* Scene.resolution = (height, width), core.py:397 */
    int32_t row_begin, row_end; /* screen-row band [row_begin,row_end) in BUFFER rows (pre-flip); 0,height = all */
    int32_t persist_silhouette; /* 1: toggle the scene's persistent silhouette like core.py:605 (Appendix B-3);
                                   0: every view starts from an empty set (fresh-Model semantics) */
    int32_t shading;            /* B2R_SHADE_*; 0 = general_shading, what the reference's render() runs */
} b2r_frame_params;

/* Optional per-view debug planes, each NULL or n_views * height * width elements (device or host pointer as the
 * `out_on_device` flag says).  Row r = buffer row r (NOT flipped), i.e. the reference's z_buffer[row, col]. */
typedef struct b2r_debug_out {
    double* z;           /* z_buffer          core.py:590 */
    int16_t* stencil;    /* stencil_buffer    core.py:591 */
    int32_t* winner;     /* global face index that coloured the pixel, -1 = background */
    uint8_t* face_status;/* n_views * total_faces: B2R_FACE_* of pass 3 (core.py:624-636) */
    int32_t* n_silhouette; /* n_views * n_models: silhouette edges extruded for that view */
    float* frame_f32;    /* n_views*height*width*3: the float32 frame BEFORE flip and tonemap (core.py:588, buffer rows);
                            what the frustum overlay of core.py:638 is drawn on (py_numpy_renderer_b200/overlay.py) */
} b2r_debug_out;

typedef struct b2r_scene b2r_scene;

int b2r_abi_version(void);
const char* b2r_last_error(void);

#define B2R_ERR_INDEX 2

/* Create (first call) or select the context of CUDA device `device` and make it current for the calling thread.
 * Must precede everything else.  b2r_shutdown drains and releases every context (destroy the scenes first). */
int b2r_init(int device);
int b2r_shutdown(void);
int b2r_current_device(void); /* device of the calling thread's current context, -1 = none */

int b2r_scene_create(const b2r_model_desc* models, int32_t n_models,
                     const b2r_texture_desc* textures, int32_t n_textures,
                     const b2r_cubemap_desc* skybox /* nullable */,
                     b2r_scene** out_scene);
int b2r_scene_destroy(b2r_scene* scene);
int b2r_scene_reset_silhouette(b2r_scene* scene);
int64_t b2r_scene_device_bytes(const b2r_scene* scene);
/* Restore a persistent silhouette set (e.g. after the device scene was rebuilt because a Model was transformed: in the
 * reference `model.silhouette` lives on the Model and survives `model @ M`).  pairs (n,2): scene-global vertex indices
 * in stored orientation, as b2r_scene_get_silhouette returns them; pairs that are no edge of the scene are ignored. */
int b2r_scene_set_silhouette(b2r_scene* scene, const int32_t* pairs, int32_t n);
/* Read the persistent silhouette set back (`model.silhouette`, core.py:251): out_pairs (capacity,2) scene-global
 * vertex indices in stored orientation, out_model (capacity) owning model.  Returns the number of edges. */
int b2r_scene_get_silhouette(b2r_scene* scene, int32_t* out_pairs, int32_t* out_model, int32_t capacity);

/* Render n_views frames of `scene`.  out_rgb: n_views*height*width*3 uint8, final image rows (flipped,
 * tonemapped: core.py:640).  With out_on_device=0 the pointers are host memory and the call returns after the
 * copies completed; with out_on_device=1 they are device memory and the call only enqueues on the library
 * stream (use b2r_sync, or order against `b2r_stream()`). */
int b2r_render(b2r_scene* scene, const b2r_frame_params* params, const b2r_view* views, int32_t n_views,
               uint8_t* out_rgb, const b2r_debug_out* debug /* nullable */, int32_t out_on_device);
/* out_on_device == 2: host pointers, but the call returns as soon as everything is enqueued; the frames (and the
 * capacity check) are complete after b2r_wait(b2r_last_ticket()).  Only these calls take a ticket.  Up to four may be
 * in flight per context: a fifth first completes the oldest (the host blocks; its outcome is kept for its b2r_wait).
 * With out_on_device == 1, b2r_sync reports a tile-list overflow of ANY render enqueued since the previous b2r_sync. */
int64_t b2r_last_ticket(void);
int b2r_wait(int64_t ticket);
int b2r_sync(void);
void* b2r_stream(void); /* the cudaStream_t the library launches on */

/* Page-locked host memory for frame buffers (cudaHostAlloc / cudaFreeHost): a device-to-host copy into pageable memory is
 * staged by the driver at a fraction of the PCIe rate.  The Python API keeps a small pool of these behind the arrays
 * `Scene.render()` returns. */
int b2r_host_alloc(int64_t bytes, void** host_ptr);
int b2r_host_free(void* host_ptr);

/* Number of kernel launches issued by this library since init (bench.py `gpu_launches`). */
int64_t b2r_launch_count(void);

/* Stage timing of the most recent b2r_render in milliseconds (CUDA events on the library stream); valid after
 * a sync.  names/ms arrays sized >= B2R_MAX_STAGES; returns the number of stages. */
#define B2R_MAX_STAGES 64
int b2r_last_stage_ms(const char** names, float* ms);
int b2r_set_stage_timing(int enabled);

/* ---- multi-GPU output window (the reference has no multi-device path: this replaces the gather that
 * follows `Scene.render()` per rank when frames are sharded) ------------------------------------------------------
 * One process per GPU.  The assembling rank creates a device buffer and exports it (64-byte CUDA IPC handle, sent to
 * the peers by the host plumbing); every peer opens the handle and passes `base + its block offset` as `out_rgb`
 * of b2r_render(..., out_on_device = 1): the shading kernel's stores then land in the assembling GPU's HBM over
 * NVLink / NVSwitch and no collective moves the frames.  A window cannot be opened by the process that created it. */
#define B2R_WINDOW_HANDLE_BYTES 64
int b2r_window_create(int64_t bytes, void** dev_ptr, void* handle_out);
int b2r_window_open(const void* handle, void** dev_ptr);
int b2r_window_close(void* dev_ptr);      /* opened with b2r_window_open */
/* Sparse push of n_views finished frames (n_views, height, width, 3) uint8 from this device into a window block (this
 * device's own memory or a peer's, mapped with b2r_window_open), asynchronously on `stream` (a cudaStream_t).  One 32x32
 * tile at a time: a tile that holds a single colour now and held exactly that after the previous push into the same
 * block is not stored again, so a frame's constant background crosses NVLink once.  `state_dev`: n_views * ceil(height/32)
 * * (width/32) words on THIS device, zeroed before the first push into a block and kept with it.  width % 32 == 0. */
int b2r_window_push(const uint8_t* src_dev, uint8_t* dst_dev, int32_t n_views, int32_t height, int32_t width,
                    uint32_t* state_dev, void* stream);
int b2r_window_destroy(void* dev_ptr);    /* created with b2r_window_create */

/* ---- native OBJ tokenizer (host only) ---------------------------------------------------------
 * The arrays `Model.load_model` builds (core.py:257-318): vertices float32 (V,4), uv float32 (T,3), normals float32
 * (N,3), faces int32 (F,3,4) = [v, vt, vn, material slot] fan-triangulated and 0-based (-1 = absent), the `usemtl`
 * names in slot order ('\n'-separated, slot 0 = "default") and the `mtllib` file names ('\n'-separated).
 * All buffers are malloc'ed by the library and released by b2r_obj_free. */
typedef struct b2r_obj {
    float* vertices; float* uv; float* normals; int32_t* faces;
    char* slot_names; char* mtllibs;
    int32_t n_vertices, n_uv, n_normals, n_faces;
} b2r_obj;
int b2r_obj_load(const char* path, b2r_obj* out);
void b2r_obj_free(b2r_obj* obj);

#ifdef __cplusplus
}
#endif
#endif /* B2R_H_ */
