// b2r_device.cuh -- device-side types and the numerics contract helpers of the B200 frame pipeline.
//
// Every helper states which NumPy/BLAS evaluation order of the reference it reproduces (SURVEY.md A.9,
// DESIGN.md "numerics contract").  The translation unit is compiled with --fmad=false, so `a*b+c` is NEVER
// contracted: a fused multiply-add happens only where fma() is written.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b2r.h"

namespace b2r {

constexpr int TILE_W = 32;  // screen tile owned by one CTA of the raster kernel
constexpr int TILE_H = 32;
constexpr int TILE_PX = TILE_W * TILE_H;
constexpr int RASTER_THREADS = 128;

// ---- per-face static data (scene lifetime) ------------------------------------------------------------------
enum : int {
    FS_CLIP = 1,        // Model.clip
    FS_VTX_F32 = 2,     // Model.vertices is float32 (facing normal / tangent edges evaluated in float32)
    FS_UV_F32 = 4,      // uv deltas of tangent_() evaluated in float32
    FS_HAS_UV = 8,
    FS_HAS_NORMALS = 16,
    FS_NO_ZWRITE = 32,
    FS_NRM_F32 = 64,    // Model.normals is float32 (`bar @ face.normals` of gouraud / pbr stays in float32)  // Model.depth_test == False: tested against z, never writes it (triangular.py:117)
};
struct FaceStatic {  // 48 B, indices are GLOBAL (scene-level concatenated arrays)
    int v[3];
    int t[3];
    int n[3];
    int material;  // global material index
    int flags;     // FS_*
    int model;
};

// Everything the shader needs about a face in ONE record (scene lifetime), so that a pixel goes winner -> record
// instead of winner -> indices -> positions / uv / normals (one dependent gather level less).
struct ShadeStatic {  // 208 B
    double wp[3][3];      // world xyz of the three corners
    double uu[3], vv[3];  // u and v of the three corners
    double vn[3][3];      // vertex normals
    int material;
    int flags;            // FS_*
};

// The same face for the float32 lighting path (k_shade_packed<SHADE_F32>): texture addressing stays float64 (a texel
// index is a discontinuous function of u, v), everything that only feeds the lighting sum is float32, and the
// per-face constants of tangent_() (edge vectors, uv deltas: core.py:205-213) are evaluated once at scene creation in
// the dtype the reference evaluates them in.
struct __align__(16) ShadeLite {  // 176 B
    double uu[3], vv[3];
    float wp[3][3];            // world xyz of the corners
    float vn[3][3];            // vertex normals; the flat unit normal three times when the model has none
    float r0[3], r1[3];        // b - a, c - a
    float du1, du2, dv1, dv2;  // uv deltas
    int material;
    int flags;
    int pad[2];
};

struct MaterialDev {
    double Kd[3];
    double Ks255[3];  // Ks * 255 (core.py:152)
    double Ns;
    int map_Kd, map_Ks, norm;
    int ns_int;  // Ns if it is a small non-negative integer (pow by squaring), else -1
    int ns_log2; // k if Ns == 2^k (k squarings), else -1
    int pad;
    double Pm, Pr, Ka[3];  // metalness, roughness, ambient colour: pbr() only (materials.py:47-49)
    float Kdf[3], Ks255f[3];  // float32 copies for the float32 lighting path
};

struct TextureDev {
    const uchar4* texels;  // RGBX
    int height, width;
    int decode;
    int tangent;
};

// ---- per-view constants -----------------------------------------------------------------------------------------
struct SkyTri {  // one of the two full-screen triangles of fill_frame_from_skybox (cube_map.py:83-101)
    long long ax, ay, v0x, v0y, v1x, v1y;
    float d00, d01, d11, inv;
    double rays[3][3];
    int ok;  // denom != 0
    int pad;
};
struct ViewDev {
    double mvp[16], mvp_dbg[16], viewport[16], planes[24];
    double cam_pos[3];
    float cam_posf[3], pad_f;
    double zl_num, zl_sum, zl_diff;  // 2*near*far, far+near, far-near  (core.py:226-228)
    int system, backface;
    SkyTri sky[2];
};

struct LightDev {
    double position[3], direction[3], color[3], ambient[3];
    double specular_strength, constant, linear, quadratic, spot_cos_outer, spot_cos_inner;
    int type;
    int pad;
};

struct LightLite {  // float32 copy of the light for the float32 lighting path
    float position[3], direction[3], color[3], ambient[3];
    float specular_strength, constant, linear, quadratic, spot_cos_outer, spot_inv_range;
};

struct FrameDev {
    LightDev light;
    LightLite lightf;
    float background[3];
    int bg_mode;
    int H, W;
    int row_begin, row_end;
    int tiles_x, tiles_y, tile_row0;  // tile grid covering the band: tile rows [tile_row0, tile_row0 + tiles_y)
    int n_faces;                      // total faces of the scene
    int sky_size;
    int want_status;
    int full_stencil;  // stencil counts are wanted for every pixel (debug plane), not only under faces
    int shading;       // B2R_SHADE_*: general_shading, or one of the alternatives of triangular.py:174-263
    int pad2;
    int* err_flag;     // mapped pinned host word: set to 1 when a texture lookup falls outside its map (IndexError)
};

// ---- per-view, per-face raster record (written by tri_setup, read by raster + shade) ---------------------------
enum : int {
    TR_VALID = 1,
    TR_NEEDS_CLIP = 2,  // per-pixel clip test cannot be skipped
    TR_BOX_ONE = 4,     // bbox holds exactly one pixel  -> N==1 evaluation order in barycentric()
    TR_COV_ONE = 8,     // exactly one covered & unclipped pixel -> N==1 order for the z interpolation
    TR_NO_ZWRITE = 16,  // face of a Model(depth_test=False)
};
struct TriRec {  // 128 B (8-byte aligned on purpose: staged copies in shared memory sit on a 136-byte pitch)
    double ax, ay, v0x, v0y, v1x, v1y;  // screen a, b-a, c-a                          48
    double zl[3];                       // linearised z at the vertices                 24
    double d[3];                        // 1/w at the vertices                          24
    float d00, d01, d11, inv;           // float32 barycentric constants                16
    short bx0, bx1, by0, by1;           // pixel box [bx0,bx1) x [by0,by1)               8
    int flags;                          //                                               4
    int pad;                            //                                               4
};

// What binning needs of a face, 16 B instead of the 128-byte record (with a million faces per view k_bin is bound by
// the bytes it reads): written by k_tri_setup for EVERY face, flags == 0 for faces that are culled / empty / clipped.
struct __align__(16) TriBox {
    short bx0, bx1, by0, by1;
    int flags;  // TR_*
    int pad;
};

// ---- shadow volume quads -------------------------------------------------------------------------------------------
struct SilEdge {  // view independent: world-space quad (A, B, D, C) of core.py:612-621
    double q[16];
};
struct __align__(16) QuadRec {  // 256 B
    double x[B2R_MAX_POLY], y[B2R_MAX_POLY];  // projected polygon                     192
    double nx, ny, nz, D;                     // plane                                  32
    int n;                                    // vertex count, 0 = rejected
    int front;
    short bx0, bx1, by0, by1;
    int pad[4];
};

// ---- evaluation-order helpers ----------------------------------------------------------------------------------------
// matrix @ matrix rows (gemm): acc = a0*b0; acc = fma(a_k, b_k, acc)
__device__ __forceinline__ double seq3(double a0, double a1, double a2, double b0, double b1, double b2) {
    return fma(a2, b2, fma(a1, b1, a0 * b0));
}
// matrix @ vector (gemv), k = 3: fma(a2,b2, fma(a0,b0, a1*b1));  k = 2: fma(a0,b0, a1*b1)
__device__ __forceinline__ double gemv3(double a0, double a1, double a2, double b0, double b1, double b2) {
    return fma(a2, b2, fma(a0, b0, a1 * b1));
}
__device__ __forceinline__ double gemv2(double a0, double a1, double b0, double b1) { return fma(a0, b0, a1 * b1); }
__device__ __forceinline__ double seq2(double a0, double a1, double b0, double b1) { return fma(a1, b1, a0 * b0); }
__device__ __forceinline__ double dot4_seq(const double* a, const double* b) {
    return fma(a[3], b[3], fma(a[2], b[2], fma(a[1], b[1], a[0] * b[0])));
}
// row vector (4) @ 4x4 row-major matrix, gemm order
__device__ __forceinline__ void vec4_mat4(const double v[4], const double* __restrict__ M, double out[4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) out[j] = fma(v[3], M[12 + j], fma(v[2], M[8 + j], fma(v[1], M[4 + j], v[0] * M[j])));
}
// np.linalg.norm(x, 2, -1): separate squares, left-to-right adds, no FMA
__device__ __forceinline__ double norm3(double x, double y, double z) { return sqrt((x * x + y * y) + z * z); }
__device__ __forceinline__ void normalize3(double v[3]) {  // transformation.py:46-49
    double l = norm3(v[0], v[1], v[2]);
    if (l == 0) l = 1;
    v[0] /= l; v[1] /= l; v[2] /= l;
}
__device__ __forceinline__ double dot3_plain(const double a[3], const double b[3]) {
    return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2];
}
__device__ __forceinline__ double linearize_z(double depth, const ViewDev& V) {
    return V.zl_num / (V.zl_sum - depth * V.zl_diff);
}

// order preserving double -> uint64 key (NaN must be filtered by the caller; -0 is folded onto +0)
__device__ __forceinline__ unsigned long long zkey(double z) {
    const long long b = __double_as_longlong(z + 0.0);
    return (unsigned long long)b ^ ((unsigned long long)(b >> 63) | 0x8000000000000000ull);
}
__device__ __forceinline__ double zkey_decode(unsigned long long k) {
    long long b = (k & 0x8000000000000000ull) ? (long long)(k & 0x7FFFFFFFFFFFFFFFull) : (long long)~k;
    return __longlong_as_double(b);
}

// float32 barycentric coverage of integer pixel (px,py) (transformation.py:12-32, triangular.py:74-78)
__device__ __forceinline__ bool tri_bary(const TriRec& r, int px, int py, float& bu, float& bv, float& bw) {
    const double v2x = (double)px - r.ax, v2y = (double)py - r.ay;
    float d20, d21;
    if (r.flags & TR_BOX_ONE) {  // (1,2)@(2,) goes through dot: seq order
        d20 = (float)seq2(v2x, v2y, r.v0x, r.v0y);
        d21 = (float)seq2(v2x, v2y, r.v1x, r.v1y);
    } else {
        d20 = (float)gemv2(v2x, v2y, r.v0x, r.v0y);
        d21 = (float)gemv2(v2x, v2y, r.v1x, r.v1y);
    }
    bv = __fmul_rn(__fsub_rn(__fmul_rn(r.d11, d20), __fmul_rn(r.d01, d21)), r.inv);
    bw = __fmul_rn(__fsub_rn(__fmul_rn(r.d00, d21), __fmul_rn(r.d01, d20)), r.inv);
    bu = __fsub_rn(__fsub_rn(1.0f, bv), bw);
    return bu >= 0.0f && bv >= 0.0f && bw >= 0.0f;
}

// perspective-correct barycentrics (core.py:155-160); n_one selects the N==1 evaluation order of the sum
__device__ __forceinline__ void persp_bary(const TriRec& r, float bu, float bv, float bw, bool n_one, double P[3]) {
    const double b0 = (double)bu, b1 = (double)bv, b2 = (double)bw;
    const double wsum = n_one ? seq3(b0, b1, b2, r.d[0], r.d[1], r.d[2]) : gemv3(b0, b1, b2, r.d[0], r.d[1], r.d[2]);
    P[0] = b0 * r.d[0] / wsum;
    P[1] = b1 * r.d[1] / wsum;
    P[2] = b2 * r.d[2] / wsum;
}

__device__ __forceinline__ bool clip_inside(const double q[4]) {
    return (-q[3] < q[0]) && (q[0] < q[3]) && (-q[3] < q[1]) && (q[1] < q[3]) && (-q[3] < q[2]) && (q[2] < q[3]);
}

// clip-space coordinates of a face's three vertices for both cameras (triangular.py:39-40), 24 doubles:
// cc[v*4 + k] = (vertex v @ MVP)[k],  cc[12 + v*4 + k] = (vertex v @ MVP_debug)[k]
constexpr int CLIP_DOUBLES = 24;
__device__ __forceinline__ bool pixel_unclipped(const double* __restrict__ cc, const double P[3], bool n_one) {
    bool ok = true;
#pragma unroll
    for (int cam = 0; cam < 2; ++cam) {
        const double* c = cc + cam * 12;
        double q[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)  // (N,3)@(3,4) is a gemm (seq); a single row goes through gemv
            q[k] = n_one ? gemv3(P[0], P[1], P[2], c[k], c[4 + k], c[8 + k]) : seq3(P[0], P[1], P[2], c[k], c[4 + k], c[8 + k]);
        ok = ok && clip_inside(q);
    }
    return ok;
}

}  // namespace b2r
