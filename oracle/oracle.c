/* oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C, sequential CPU restatement of the reference frame pipeline `Scene.render()`
 * (Denizantip/py-numpy-renderer, obj/core.py:587-640) used ONLY as the parity checker for the CUDA path
 * (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference legs).  The product never links,
 * imports or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this file against fixtures in tests/golden/ that
 * were produced by running the unmodified Python reference in the build container (oracle/make_golden.py):
 * z-buffer, stencil counts and per-pixel winner face bit-exact, RGB within 1 LSB.
 *
 * It deliberately keeps the reference's STRUCTURE (three sequential passes over faces in array order, an
 * in-place z-buffer, a toggled silhouette set) -- the CUDA path uses an order-independent formulation, so the
 * two are independent statements of the same semantics.
 *
 * Floating-point evaluation order follows what NumPy 2.3.5 / OpenBLAS 0.3.30 do for each call site (SURVEY.md
 * A.9, re-probed with exact rational arithmetic -- see DESIGN.md "numerics contract"):
 *   (N>=2,k)@(k,m) matrix products : acc = a0*b0; acc = fma(a_k, b_k, acc)                      ("seq")
 *   (N>=2,k)@(k,) and (N>=2,k)@(k,1): k=3 fma(a2,b2,fma(a0,b0,a1*b1)); k=2 fma(a0,b0,a1*b1)     ("gemv")
 *   N==1 vector.vector              : seq                   N==1 (1,k)@(k,m>1): gemv order
 *   1-D . 1-D                       : seq
 *   np.cross, np.linalg.norm(x,2,-1), (x*y).sum(-1): separate mul / add, left to right, no FMA
 *   float32 scalar arithmetic in `barycentric`: every op rounded to float32
 * Build: gcc -O2 -ffp-contract=off -mfma -pthread (see oracle/Makefile); -ffp-contract=off is REQUIRED so that
 * only the explicit fma() calls fuse.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/b2r.h"

#if defined(__FP_FAST_FMA) || defined(__FMA__)
#define FMA(a, b, c) __builtin_fma((a), (b), (c))
#else
#define FMA(a, b, c) fma((a), (b), (c))
#endif

/* ---------------------------------------------------------------------------------------------------------- */
/* evaluation-order helpers                                                                                   */
static inline double dot_seq(const double* a, const double* b, int n) {
    double acc = a[0] * b[0];
    for (int k = 1; k < n; ++k) acc = FMA(a[k], b[k], acc);
    return acc;
}
static inline double gemv3(double a0, double a1, double a2, double b0, double b1, double b2) {
    return FMA(a2, b2, FMA(a0, b0, a1 * b1));
}
static inline double seq3(double a0, double a1, double a2, double b0, double b1, double b2) {
    return FMA(a2, b2, FMA(a1, b1, a0 * b0));
}
static inline double gemv2(double a0, double a1, double b0, double b1) { return FMA(a0, b0, a1 * b1); }
static inline double seq2(double a0, double a1, double b0, double b1) { return FMA(a1, b1, a0 * b0); }
/* (N,3)@(3,) with the N==1 special case */
static inline double vec3(int n_one, const double* a, double b0, double b1, double b2) {
    return n_one ? seq3(a[0], a[1], a[2], b0, b1, b2) : gemv3(a[0], a[1], a[2], b0, b1, b2);
}
/* (N,3)@(3,m>1) with the N==1 special case (vector @ matrix goes through gemv) */
static inline double mat3(int n_one, const double* a, double b0, double b1, double b2) {
    return n_one ? gemv3(a[0], a[1], a[2], b0, b1, b2) : seq3(a[0], a[1], a[2], b0, b1, b2);
}
static inline void vec4_mat4(const double v[4], const double M[16], double out[4]) {
    for (int j = 0; j < 4; ++j) {
        double acc = v[0] * M[j];
        acc = FMA(v[1], M[4 + j], acc);
        acc = FMA(v[2], M[8 + j], acc);
        acc = FMA(v[3], M[12 + j], acc);
        out[j] = acc;
    }
}
static inline void cross3(const double a[3], const double b[3], double c[3]) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}
static inline double norm3(const double a[3]) { return sqrt((a[0] * a[0] + a[1] * a[1]) + a[2] * a[2]); }
/* transformation.py:46-49 */
static inline void normalize3(const double a[3], double out[3]) {
    double l = norm3(a);
    if (l == 0) l = 1;
    out[0] = a[0] / l; out[1] = a[1] / l; out[2] = a[2] / l;
}
static inline double dot3_nofma(const double a[3], const double b[3]) {
    return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2];
}

/* ---------------------------------------------------------------------------------------------------------- */
typedef struct {
    const b2r_model_desc* models; int n_models;
    const b2r_texture_desc* textures; int n_textures;
    const b2r_cubemap_desc* sky;
    const b2r_frame_params* fp;
    const b2r_view* view;
    int H, W;
    float* frame;     /* (H,W,3) */
    double* z;        /* (H,W)   */
    int16_t* stencil; /* (H,W)   */
    int32_t* winner1; /* (H,W) last face that coloured the pixel in pass 1 */
    int32_t* winner3; /* (H,W) last face that coloured the pixel in pass 3 */
    float lut_unorm[256], lut_snorm[256];
    int index_error;  /* a texture lookup fell outside its map: the reference raises IndexError (core.py:138-173) */
} ctx_t;

static inline double load_real(const void* base, int dtype, int64_t idx) {
    return dtype == B2R_F32 ? (double)((const float*)base)[idx] : ((const double*)base)[idx];
}

typedef struct {
    const b2r_model_desc* m;
    int32_t vi[3], ti[3], ni[3], slot;
    double wv[3][4]; /* world vertices (exact promotion) */
    const b2r_material* mat;
} face_t;

static void face_load(const b2r_model_desc* m, int f, face_t* F) {
    F->m = m;
    const int32_t* r = m->faces + (int64_t)f * 12;
    for (int c = 0; c < 3; ++c) {
        F->vi[c] = r[c * 4 + 0]; F->ti[c] = r[c * 4 + 1]; F->ni[c] = r[c * 4 + 2];
        int32_t v = F->vi[c];
        if (v < 0) v += m->n_vertices; /* python negative index */
        for (int k = 0; k < 4; ++k) F->wv[c][k] = load_real(m->vertices, m->vertex_dtype, (int64_t)v * 4 + k);
    }
    int32_t slot = r[3];
    if (slot < 0) slot += m->n_materials;
    if (slot < 0 || slot >= m->n_materials) slot = 0;
    F->slot = slot;
    F->mat = m->materials + slot;
}

/* Face.unit_normal_world_space (core.py:127-130): arithmetic in the vertex dtype. */
static void unit_normal_world(const face_t* F, double n[3]) {
    if (F->m->vertex_dtype == B2R_F32) {
        float a[3], b[3], c[3], e0[3], e1[3], cr[3];
        for (int k = 0; k < 3; ++k) { a[k] = (float)F->wv[0][k]; b[k] = (float)F->wv[1][k]; c[k] = (float)F->wv[2][k]; }
        for (int k = 0; k < 3; ++k) { e0[k] = b[k] - a[k]; e1[k] = c[k] - a[k]; }
        cr[0] = e0[1] * e1[2] - e0[2] * e1[1];
        cr[1] = e0[2] * e1[0] - e0[0] * e1[2];
        cr[2] = e0[0] * e1[1] - e0[1] * e1[0];
        float l = sqrtf((cr[0] * cr[0] + cr[1] * cr[1]) + cr[2] * cr[2]);
        if (l == 0) l = 1;
        for (int k = 0; k < 3; ++k) n[k] = (double)(cr[k] / l);
    } else {
        double e0[3], e1[3], cr[3];
        for (int k = 0; k < 3; ++k) { e0[k] = F->wv[1][k] - F->wv[0][k]; e1[k] = F->wv[2][k] - F->wv[0][k]; }
        cross3(e0, e1, cr);
        normalize3(cr, n);
    }
}

/* Face.linearize_z (core.py:226-228) */
static inline double linearize_z(double depth, double near_, double far_) {
    return (2 * near_ * far_) / (far_ + near_ - depth * (far_ - near_));
}

/* Face.get_UV (core.py:138-143): returns row, col with python negative-index wrap; -1 row => invalid */
static void get_uv_texel(const ctx_t* C, const b2r_texture_desc* T, const double P[3], const double uvu[3],
                         const double uvv[3], int n_one, int* row, int* col) {
    double cu = vec3(n_one, P, uvu[0], uvu[1], uvu[2]);
    double cv = vec3(n_one, P, uvv[0], uvv[1], uvv[2]);
    if (cu > 1.0) cu = 1.0;
    double rv = 1.0 - cv;
    if (rv > 1.0) rv = 1.0;
    int32_t c = (int32_t)(cu * (T->width - 1));
    int32_t r = (int32_t)(rv * (T->height - 1));
    if (c < 0) c += T->width;
    if (r < 0) r += T->height;
    /* below -size NumPy's fancy indexing raises IndexError (NaN: astype(int32) gives INT_MIN): remembered, the
     * render still completes so that the planes can be inspected */
    if (c < 0 || c >= T->width || cu != cu) { c = 0; ((ctx_t*)C)->index_error = 1; }
    if (r < 0 || r >= T->height || rv != rv) { r = 0; ((ctx_t*)C)->index_error = 1; }
    *row = r; *col = c;
}
static inline void texel_f32(const ctx_t* C, const b2r_texture_desc* T, int row, int col, float out[3]) {
    const uint8_t* p = T->rgb + ((int64_t)row * T->width + col) * 3;
    const float* lut = T->decode == B2R_TEX_SNORM ? C->lut_snorm : C->lut_unorm;
    out[0] = lut[p[0]]; out[1] = lut[p[1]]; out[2] = lut[p[2]];
}

/* np.linalg.inv of a 3x3 (core.py:214): LU with partial pivoting (what LAPACK dgesv does), then the solves
 * against the identity.  Not bit-reproducible against LAPACK -- covered by the RGB tolerance. */
static void inv3(const double A[9], double out[9]) {
    double a[9]; int perm[3] = {0, 1, 2};
    memcpy(a, A, sizeof(a));
    for (int k = 0; k < 3; ++k) {
        int p = k; double best = fabs(a[k * 3 + k]);
        for (int r = k + 1; r < 3; ++r) if (fabs(a[r * 3 + k]) > best) { best = fabs(a[r * 3 + k]); p = r; }
        if (p != k) {
            for (int c = 0; c < 3; ++c) { double t = a[k * 3 + c]; a[k * 3 + c] = a[p * 3 + c]; a[p * 3 + c] = t; }
            int t = perm[k]; perm[k] = perm[p]; perm[p] = t;
        }
        double piv = 1.0 / a[k * 3 + k];
        for (int r = k + 1; r < 3; ++r) {
            a[r * 3 + k] *= piv;
            for (int c = k + 1; c < 3; ++c) a[r * 3 + c] -= a[r * 3 + k] * a[k * 3 + c];
        }
    }
    for (int col = 0; col < 3; ++col) {
        double y[3];
        for (int r = 0; r < 3; ++r) y[r] = perm[r] == col ? 1.0 : 0.0;
        for (int r = 1; r < 3; ++r) for (int c = 0; c < r; ++c) y[r] -= a[r * 3 + c] * y[c];
        for (int r = 2; r >= 0; --r) {
            for (int c = r + 1; c < 3; ++c) y[r] -= a[r * 3 + c] * y[c];
            y[r] /= a[r * 3 + r];
        }
        for (int r = 0; r < 3; ++r) out[r * 3 + col] = y[r];
    }
}

/* The alternative shading functions whose call sites are commented out in rasterize() (triangular.py:120-130):
 * flat_shading (174-177), gouraud (180-182), pbr (222-263).  `sv` = the face's SCREEN vertices as rasterize() leaves them
 * in face.vertices when the shader runs (viewport x, y, linearised z): pbr() uses them as positions.  Values are written
 * to the float32 frame like the reference does (flat / gouraud: up to 255, the tonemap then wraps in uint8). */
static inline float clampf_d(double v, double lo, double hi) { return (float)(v < lo ? lo : (v > hi ? hi : v)); }
static void normalize3f(const float a[3], float out[3]) { /* normalize() on a float32 array (transformation.py:46-49) */
    float l = sqrtf((a[0] * a[0] + a[1] * a[1]) + a[2] * a[2]);
    if (l == 0) l = 1;
    out[0] = a[0] / l; out[1] = a[1] / l; out[2] = a[2] / l;
}
static void shade_alt(const ctx_t* C, const face_t* F, const float bar[3], const double sv[3][3], int n_one, float out[3]) {
    const b2r_light* L = &C->fp->light;
    const b2r_model_desc* m = F->m;
    const int mode = C->fp->shading;
    if (mode == B2R_SHADE_FLAT) {
        double n[3];
        unit_normal_world(F, n);
        double it = dot_seq(n, L->direction, 3);          /* face.unit_normal_world_space @ light.direction */
        it = it < 0.3 ? 0.3 : (it > 1.0 ? 1.0 : it);     /* NaN stays NaN through both comparisons */
        out[0] = out[1] = out[2] = (float)(it * 255);
        return;
    }
    double vn[3][3] = {{0}};
    if (m->normals) for (int c = 0; c < 3; ++c) {
        int32_t t = F->ni[c]; if (t < 0) t += m->n_normals;
        for (int k = 0; k < 3; ++k) vn[c][k] = load_real(m->normals, m->normal_dtype, (int64_t)t * 3 + k);
    }
    /* bar @ face.normals: float32 @ float32 stays float32 (sgemm), float32 @ float64 is evaluated in float64 */
    double nb[3];
    if (m->normal_dtype == B2R_F32) {
        for (int k = 0; k < 3; ++k) {
            float acc = n_one ? fmaf(bar[2], (float)vn[2][k], fmaf(bar[0], (float)vn[0][k], bar[1] * (float)vn[1][k]))
                              : fmaf(bar[2], (float)vn[2][k], fmaf(bar[1], (float)vn[1][k], bar[0] * (float)vn[0][k]));
            nb[k] = (double)acc;
        }
    } else {
        double b[3] = {(double)bar[0], (double)bar[1], (double)bar[2]};
        for (int k = 0; k < 3; ++k) nb[k] = mat3(n_one, b, vn[0][k], vn[1][k], vn[2][k]);
    }
    if (mode == B2R_SHADE_GOURAUD) {
        double it = (nb[0] * L->direction[0] + nb[1] * L->direction[1]) + nb[2] * L->direction[2];
        it = it < 0.0 ? 0.0 : (it > 1.0 ? 1.0 : it);
        out[0] = out[1] = out[2] = (float)(it * 255);
        return;
    }
    /* pbr (triangular.py:222-263) */
    const double PI = 3.141592653589793;
    const double metallic = F->mat->Pm, roughness = F->mat->Pr;
    double N[3];
    if (m->normal_dtype == B2R_F32) {
        float nf[3] = {(float)nb[0], (float)nb[1], (float)nb[2]}, nn[3];
        normalize3f(nf, nn);
        N[0] = nn[0]; N[1] = nn[1]; N[2] = nn[2];
    } else normalize3(nb, N);
    double b[3] = {(double)bar[0], (double)bar[1], (double)bar[2]}, pos[3];
    for (int k = 0; k < 3; ++k) pos[k] = mat3(n_one, b, sv[0][k], sv[1][k], sv[2][k]);
    double vv[3] = {C->view->cam_pos[0] - pos[0], C->view->cam_pos[1] - pos[1], C->view->cam_pos[2] - pos[2]}, Vd[3];
    normalize3(vv, Vd);
    const double F0 = 0.04 * (1 - metallic) + 1.0 * metallic;       /* mix(F0, albedo = 1, metallic) */
    double lv[3] = {L->position[0] - pos[0], L->position[1] - pos[1], L->position[2] - pos[2]}, Ld[3];
    normalize3(lv, Ld);
    double hv[3] = {Vd[0] + Ld[0], Vd[1] + Ld[1], Vd[2] + Ld[2]}, Hd[3];
    normalize3(hv, Hd);
    const double distance = norm3(lv);
    const double attenuation = 1.0 / (distance * distance);
    /* DistributionGGX */
    const double a = roughness * roughness, a2 = a * a;
    double NdotH = dot3_nofma(N, Hd); if (NdotH < 0) NdotH = 0;
    double den = NdotH * NdotH * (a2 - 1.0) + 1.0;
    den = PI * den * den;
    const double NDF = a2 / den;
    /* GeometrySmith */
    double NdotV = dot3_nofma(N, Vd); if (NdotV < 0) NdotV = 0;
    double NdotL = dot3_nofma(N, Ld); if (NdotL < 0) NdotL = 0;
    const double r1 = roughness + 1.0, kk = (r1 * r1) / 8.0;
    const double G = (NdotL / (NdotL * (1.0 - kk) + kk)) * (NdotV / (NdotV * (1.0 - kk) + kk));
    double HdotV = dot3_nofma(Hd, Vd); if (HdotV < 0) HdotV = 0;
    const double Fr = F0 + (1.0 - F0) * pow(1 - HdotV, 5);
    const double kD = (1.0 - Fr) * (1.0 - metallic);
    const double specular = (NDF * G * Fr) / (4.0 * NdotV * NdotL + 0.0001);
    for (int k = 0; k < 3; ++k) {
        const double radiance = L->color[k] * attenuation;
        const double Lo = (kD * 1.0 / PI + specular) * radiance * NdotL;
        double color = 1.0 * F->mat->Ka[k] + Lo;
        color = color / (color + 1.0);
        out[k] = (float)pow(color, 1.0 / 2.2);
    }
}

/* general_shading (triangular.py:135-171) for one pixel; bar = screen barycentrics (float32), d = 1/w per vertex */
static void shade_pixel(const ctx_t* C, const face_t* F, const float bar[3], const double d[3], int first_pass,
                        int n_one, float out[3]) {
    const b2r_light* L = &C->fp->light;
    const b2r_model_desc* m = F->m;
    /* Face.screen_perspective (core.py:155-160) */
    double b[3] = {(double)bar[0], (double)bar[1], (double)bar[2]};
    double wsum = vec3(n_one, b, d[0], d[1], d[2]);
    double P[3] = {b[0] * d[0] / wsum, b[1] * d[1] / wsum, b[2] * d[2] / wsum};
    double uvu[3] = {0, 0, 0}, uvv[3] = {0, 0, 0};
    if (m->uv) for (int c = 0; c < 3; ++c) {
        int32_t t = F->ti[c]; if (t < 0) t += m->n_uv;
        uvu[c] = load_real(m->uv, m->uv_dtype, (int64_t)t * 3 + 0);
        uvv[c] = load_real(m->uv, m->uv_dtype, (int64_t)t * 3 + 1);
    }
    /* Face.get_object_color (core.py:162-173) */
    double albedo[3];
    if (F->mat->map_Kd >= 0) {
        const b2r_texture_desc* T = C->textures + F->mat->map_Kd;
        int r, c; float t[3];
        get_uv_texel(C, T, P, uvu, uvv, n_one, &r, &c);
        texel_f32(C, T, r, c, t);
        for (int k = 0; k < 3; ++k) albedo[k] = (double)t[k];
    } else {
        for (int k = 0; k < 3; ++k) albedo[k] = F->mat->Kd[k];
    }
    double frag[3];
    for (int k = 0; k < 3; ++k) frag[k] = mat3(n_one, P, F->wv[0][k], F->wv[1][k], F->wv[2][k]);
    /* Light.attenuation (core.py:517-524) */
    double dl[3] = {L->position[0] - frag[0], L->position[1] - frag[1], L->position[2] - frag[2]};
    double dist = norm3(dl);
    double att = 1.0 / (L->constant + dist * (L->linear + L->quadratic * dist));
    if (first_pass) {
        for (int k = 0; k < 3; ++k) {
            double v = att * L->ambient[k] * albedo[k];
            v = v < 0.05 ? 0.05 : (v > 1.0 ? 1.0 : v);
            out[k] = (float)v;
        }
        return;
    }
    /* Face.get_normals (core.py:175-189) */
    double nrm[3];
    double vn[3][3];
    if (m->normals) for (int c = 0; c < 3; ++c) {
        int32_t t = F->ni[c]; if (t < 0) t += m->n_normals;
        for (int k = 0; k < 3; ++k) vn[c][k] = load_real(m->normals, m->normal_dtype, (int64_t)t * 3 + k);
    }
    if (F->mat->norm >= 0) {
        const b2r_texture_desc* T = C->textures + F->mat->norm;
        int r, c; float t[3];
        get_uv_texel(C, T, P, uvu, uvv, n_one, &r, &c);
        texel_f32(C, T, r, c, t);
        if (T->tangent) {
            /* Face.tangent_ (core.py:191-224) */
            double nn[3], n[3];
            for (int k = 0; k < 3; ++k) nn[k] = mat3(n_one, P, vn[0][k], vn[1][k], vn[2][k]);
            normalize3(nn, n);
            double A[9], AI[9];
            if (m->vertex_dtype == B2R_F32) {
                for (int k = 0; k < 3; ++k) {
                    A[0 + k] = (double)((float)F->wv[1][k] - (float)F->wv[0][k]);
                    A[3 + k] = (double)((float)F->wv[2][k] - (float)F->wv[0][k]);
                }
            } else {
                for (int k = 0; k < 3; ++k) { A[0 + k] = F->wv[1][k] - F->wv[0][k]; A[3 + k] = F->wv[2][k] - F->wv[0][k]; }
            }
            for (int k = 0; k < 3; ++k) A[6 + k] = n[k];
            inv3(A, AI);
            double du1, du2, dv1, dv2;
            if (m->uv_dtype == B2R_F32) {
                du1 = (double)((float)uvu[1] - (float)uvu[0]); du2 = (double)((float)uvu[2] - (float)uvu[0]);
                dv1 = (double)((float)uvv[1] - (float)uvv[0]); dv2 = (double)((float)uvv[2] - (float)uvv[0]);
            } else {
                du1 = uvu[1] - uvu[0]; du2 = uvu[2] - uvu[0]; dv1 = uvv[1] - uvv[0]; dv2 = uvv[2] - uvv[0];
            }
            double ti[3], tj[3], ui[3], uj[3];
            for (int r2 = 0; r2 < 3; ++r2) {
                ti[r2] = gemv3(AI[r2 * 3 + 0], AI[r2 * 3 + 1], AI[r2 * 3 + 2], du1, du2, 0.0);
                tj[r2] = gemv3(AI[r2 * 3 + 0], AI[r2 * 3 + 1], AI[r2 * 3 + 2], dv1, dv2, 0.0);
            }
            normalize3(ti, ui); normalize3(tj, uj);
            double tx[3] = {(double)t[0], (double)t[1], (double)t[2]};
            for (int r2 = 0; r2 < 3; ++r2) nrm[r2] = seq3(ui[r2], uj[r2], n[r2], tx[0], tx[1], tx[2]);
        } else {
            for (int k = 0; k < 3; ++k) nrm[k] = (double)t[k];
        }
    } else if (m->normals) {
        for (int k = 0; k < 3; ++k) nrm[k] = mat3(n_one, P, vn[0][k], vn[1][k], vn[2][k]);
    } else {
        double fn[3];
        unit_normal_world(F, fn);
        for (int k = 0; k < 3; ++k) nrm[k] = mat3(n_one, P, fn[k], fn[k], fn[k]);
    }
    double N[3]; normalize3(nrm, N);
    double Ld[3];
    if (L->type == B2R_LIGHT_DIRECTIONAL) { Ld[0] = L->direction[0]; Ld[1] = L->direction[1]; Ld[2] = L->direction[2]; }
    else normalize3(dl, Ld);
    double dv[3] = {C->view->cam_pos[0] - frag[0], C->view->cam_pos[1] - frag[1], C->view->cam_pos[2] - frag[2]};
    double Vd[3]; normalize3(dv, Vd);
    if (L->type == B2R_LIGHT_SPOT) {
        double x = dot3_nofma(L->direction, Ld);
        x = (x - L->spot_cos_outer) / (L->spot_cos_inner - L->spot_cos_outer);
        x = x < 0.0 ? 0.0 : (x > 1.0 ? 1.0 : x);
        double s = x * x * (3 - 2 * x);
        for (int k = 0; k < 3; ++k) albedo[k] = albedo[k] * s;
    }
    /* Face.get_specular (core.py:145-153) */
    double spec_light[3];
    if (F->mat->map_Ks >= 0) {
        const b2r_texture_desc* T = C->textures + F->mat->map_Ks;
        int r, c; float t[3];
        get_uv_texel(C, T, P, uvu, uvv, n_one, &r, &c);
        texel_f32(C, T, r, c, t);
        float s = t[0] * 255.0f;
        spec_light[0] = spec_light[1] = spec_light[2] = (double)s;
    } else {
        for (int k = 0; k < 3; ++k) spec_light[k] = F->mat->Ks[k] * 255;
    }
    double hv[3] = {Ld[0] + Vd[0], Ld[1] + Vd[1], Ld[2] + Vd[2]};
    double Hd[3]; normalize3(hv, Hd);
    double nh = dot3_nofma(N, Hd);
    if (nh < 0) nh = 0;
    double sr = pow(nh, F->mat->Ns);
    double nl = dot3_nofma(N, Ld);
    for (int k = 0; k < 3; ++k) {
        double specular = L->color[k] * sr * L->specular_strength * spec_light[k];
        double diffuse = nl * L->color[k];
        double v = att * albedo[k] * (L->ambient[k] + diffuse + specular);
        v = v < 0.05 ? 0.05 : (v > 1.0 ? 1.0 : v);
        out[k] = (float)v;
    }
}

static inline int clip_inside(const double q[4]) {
    return (-q[3] < q[0]) && (q[0] < q[3]) && (-q[3] < q[1]) && (q[1] < q[3]) && (-q[3] < q[2]) && (q[2] < q[3]);
}

/* transformation.py:35-43 on `count` points with stride 4 doubles; returns 0 if None */
static int bound_box(const double* pts, int count, int H, int W, int32_t box[4]) {
    double mnx = pts[0], mxx = pts[0], mny = pts[1], mxy = pts[1];
    for (int i = 1; i < count; ++i) {
        double x = pts[i * 4], y = pts[i * 4 + 1];
        if (x < mnx) mnx = x; if (x > mxx) mxx = x;
        if (y < mny) mny = y; if (y > mxy) mxy = y;
    }
    if (mnx < 0) mnx = 0; if (mxx > W) mxx = W;
    if (mny < 0) mny = 0; if (mxy > H) mxy = H;
    if (mnx > mxx || mny > mxy) return 0;
    box[0] = (int32_t)ceil(mnx); box[1] = (int32_t)ceil(mxx); box[2] = (int32_t)ceil(mny); box[3] = (int32_t)ceil(mxy);
    return 1;
}

/* rasterize (triangular.py:29-132).  stencil_pass = 0: pass 1, 1: pass 3.  Returns a B2R_FACE_* status. */
static int rasterize(ctx_t* C, const face_t* F, int face_global, int stencil_pass) {
    const b2r_view* V = C->view;
    const int H = C->H, W = C->W;
    double cs[3][4], csd[3][4], v[3][4], depth[3];
    for (int i = 0; i < 3; ++i) {
        vec4_mat4(F->wv[i], V->mvp, cs[i]);
        vec4_mat4(F->wv[i], V->mvp_dbg, csd[i]);
        depth[i] = 1 / cs[i][3];
        double t[4];
        for (int k = 0; k < 4; ++k) t[k] = cs[i][k] * depth[i];
        vec4_mat4(t, V->viewport, v[i]);
        v[i][3] = depth[i];
    }
    if (V->backface_culling) {
        double e0[3], e1[3], cr[3], n[3];
        for (int k = 0; k < 3; ++k) { e0[k] = v[1][k] - v[0][k]; e1[k] = v[2][k] - v[0][k]; }
        cross3(e0, e1, cr);
        normalize3(cr, n);
        if (n[2] < 0) return B2R_FACE_BACK_FACE_CULLING;
    }
    int32_t box[4];
    if (!bound_box(&v[0][0], 3, H, W, box)) return B2R_FACE_EMPTY_Z;
    const int nx = box[1] > box[0] ? box[1] - box[0] : 0, ny = box[3] > box[2] ? box[3] - box[2] : 0;
    const int64_t n_box = (int64_t)nx * ny;
    /* barycentric (transformation.py:12-32) */
    double v0[2] = {v[1][0] - v[0][0], v[1][1] - v[0][1]}, v1[2] = {v[2][0] - v[0][0], v[2][1] - v[0][1]};
    float d00 = (float)seq2(v0[0], v0[1], v0[0], v0[1]);
    float d01 = (float)seq2(v0[0], v0[1], v1[0], v1[1]);
    float d11 = (float)seq2(v1[0], v1[1], v1[0], v1[1]);
    float denom = d00 * d11 - d01 * d01;
    if (denom == 0) return B2R_FACE_EMPTY_B;
    float inv = 1.0f / denom;
    if (n_box == 0) return B2R_FACE_CLIPPED;
    const int box_one = n_box == 1;
    double zl[3];
    for (int i = 0; i < 3; ++i) zl[i] = linearize_z(v[i][2], V->near_, V->far_);

    /* pass A over the box: coverage & clip, count survivors (N of `bar_screen[Bi]`) */
    int64_t n_cov = 0;
    uint8_t* mask = (uint8_t*)malloc((size_t)n_box);
    float* bars = (float*)malloc((size_t)n_box * 3 * sizeof(float));
    for (int ix = 0; ix < nx; ++ix) for (int iy = 0; iy < ny; ++iy) {
        const int64_t o = (int64_t)ix * ny + iy;
        const int px = box[0] + ix, py = box[2] + iy;
        double v2[2] = {(double)px - v[0][0], (double)py - v[0][1]};
        float d20 = (float)(box_one ? seq2(v2[0], v2[1], v0[0], v0[1]) : gemv2(v2[0], v2[1], v0[0], v0[1]));
        float d21 = (float)(box_one ? seq2(v2[0], v2[1], v1[0], v1[1]) : gemv2(v2[0], v2[1], v1[0], v1[1]));
        float bv = (d11 * d20 - d01 * d21) * inv;
        float bw = (d00 * d21 - d01 * d20) * inv;
        float bu = 1.0f - bv - bw;
        bars[o * 3] = bu; bars[o * 3 + 1] = bv; bars[o * 3 + 2] = bw;
        int in = (bu >= 0) && (bv >= 0) && (bw >= 0);
        if (F->m->clip) {
            double b[3] = {(double)bu, (double)bv, (double)bw};
            double wsum = vec3(box_one, b, depth[0], depth[1], depth[2]);
            double P[3] = {b[0] * depth[0] / wsum, b[1] * depth[1] / wsum, b[2] * depth[2] / wsum};
            double q[4], qd[4];
            for (int k = 0; k < 4; ++k) {
                q[k] = mat3(box_one, P, cs[0][k], cs[1][k], cs[2][k]);
                qd[k] = mat3(box_one, P, csd[0][k], csd[1][k], csd[2][k]);
            }
            in = in && clip_inside(q) && clip_inside(qd);
        }
        mask[o] = (uint8_t)in;
        n_cov += in;
    }
    int status = B2R_FACE_RENDERED;
    if (n_cov == 0) { status = B2R_FACE_CLIPPED; goto done; }
    {
        const int cov_one = n_cov == 1;
        /* z test (triangular.py:96-112) */
        int64_t n_pass = 0;
        double* zs = (double*)malloc((size_t)n_box * sizeof(double));
        int any_z = 0;
        for (int64_t o = 0; o < n_box; ++o) {
            if (!mask[o]) continue;
            const int px = box[0] + (int)(o / ny), py = box[2] + (int)(o % ny);
            double b[3] = {(double)bars[o * 3], (double)bars[o * 3 + 1], (double)bars[o * 3 + 2]};
            double z = vec3(cov_one, b, zl[0], zl[1], zl[2]);
            zs[o] = z;
            const double zb = C->z[(int64_t)py * W + px];
            int pass = V->system == 1 ? (zb >= z) : (zb <= z);
            any_z |= pass;
            if (pass && stencil_pass) pass = C->stencil[(int64_t)py * W + px] == 0;
            mask[o] = (uint8_t)(pass ? 2 : 0);
            n_pass += pass;
        }
        (void)any_z;
        if (n_pass == 0) { status = B2R_FACE_EMPTY_Z; free(zs); goto done; }
        const int pass_one = n_pass == 1;
        for (int64_t o = 0; o < n_box; ++o) {
            if (mask[o] != 2) continue;
            const int px = box[0] + (int)(o / ny), py = box[2] + (int)(o % ny);
            const int64_t pix = (int64_t)py * W + px;
            if (!stencil_pass && F->m->depth_test) C->z[pix] = zs[o];
            if (C->fp->shading != B2R_SHADE_GENERAL) {
                const double sv[3][3] = {{v[0][0], v[0][1], zl[0]}, {v[1][0], v[1][1], zl[1]}, {v[2][0], v[2][1], zl[2]}};
                shade_alt(C, F, bars + o * 3, sv, pass_one, C->frame + pix * 3);
            } else
                shade_pixel(C, F, bars + o * 3, depth, !stencil_pass, pass_one, C->frame + pix * 3);
            (stencil_pass ? C->winner3 : C->winner1)[pix] = face_global;
        }
        free(zs);
    }
done:
    free(mask); free(bars);
    return status;
}

/* ---- shadow volumes ---------------------------------------------------------------------------------------- */
typedef struct { int32_t a, b; } edge_t;
typedef struct { edge_t* e; uint8_t* used; int cap, count; } edgeset_t;

static uint64_t edge_hash(int32_t a, int32_t b) {
    uint32_t lo = a < b ? (uint32_t)a : (uint32_t)b, hi = a < b ? (uint32_t)b : (uint32_t)a;
    uint64_t h = ((uint64_t)hi << 32 | lo) * 0x9E3779B97F4A7C15ull;
    return h ^ (h >> 29);
}
static void edgeset_init(edgeset_t* s, int cap_pow2) {
    s->cap = cap_pow2; s->count = 0;
    s->e = (edge_t*)calloc((size_t)cap_pow2, sizeof(edge_t));
    s->used = (uint8_t*)calloc((size_t)cap_pow2, 1); /* 0 empty, 1 live, 2 tombstone */
}
static void edgeset_free(edgeset_t* s) { free(s->e); free(s->used); }
/* shadow_volumes' set toggle with Edge's undirected equality (triangular.py:286-302) */
static void edgeset_toggle(edgeset_t* s, int32_t a, int32_t b) {
    uint64_t h = edge_hash(a, b) & (uint64_t)(s->cap - 1);
    int64_t first_free = -1;
    for (;;) {
        if (s->used[h] == 0) break;
        if (s->used[h] == 1) {
            edge_t* e = &s->e[h];
            if ((e->a == a && e->b == b) || (e->a == b && e->b == a)) { s->used[h] = 2; s->count--; return; }
        } else if (first_free < 0) first_free = (int64_t)h;
        h = (h + 1) & (uint64_t)(s->cap - 1);
    }
    if (first_free >= 0) h = (uint64_t)first_free;
    s->e[h].a = a; s->e[h].b = b; s->used[h] = 1; s->count++;
}

/* clipping + helpers (plane_intersection.py:24-40, 59-86) */
static int clip_polygon(const double* in, int n_in, const double planes[24], double* out) {
    double bufA[B2R_MAX_POLY * 2][4], bufB[B2R_MAX_POLY * 2][4];
    double (*cur)[4] = bufA; double (*nxt)[4] = bufB;
    int n = n_in;
    memcpy(cur, in, sizeof(double) * 4 * (size_t)n_in);
    for (int p = 0; p < 6; ++p) {
        const double* pl = planes + p * 4;
        int m = 0;
        for (int i = 0; i < n; ++i) {
            const double* c = cur[i];
            const double* nx = cur[(i + 1) % n];
            const int cv = dot_seq(pl, c, 4) >= 0, nv = dot_seq(pl, nx, 4) >= 0;
            if (cv) { memcpy(nxt[m++], c, sizeof(double) * 4); }
            if (cv ^ nv) {
                /* line_plane_intersection(next, current, plane) */
                double dir[4]; for (int k = 0; k < 4; ++k) dir[k] = c[k] - nx[k];
                double den = dot_seq(pl, dir, 4);
                if (!(fabs(den) < 1e-10)) {
                    double wgt = -dot_seq(pl, nx, 4) / den;
                    if (0 <= wgt && wgt <= 1) { for (int k = 0; k < 4; ++k) nxt[m][k] = nx[k] + wgt * dir[k]; m++; }
                }
            }
        }
        n = m;
        double (*t)[4] = cur; cur = nxt; nxt = t;
        if (n == 0) break;
    }
    memcpy(out, cur, sizeof(double) * 4 * (size_t)n);
    return n;
}

/* resterize_quadrangle (triangular.py:319-368) */
static void raster_quad(ctx_t* C, const double quad_world[16]) {
    const b2r_view* V = C->view;
    const int H = C->H, W = C->W;
    double poly[B2R_MAX_POLY * 2][4];
    int n = clip_polygon(quad_world, 4, V->planes, &poly[0][0]);
    if (n < 3) return;
    double scr[B2R_MAX_POLY * 2][4];
    for (int i = 0; i < n; ++i) {
        double c[4], t[4];
        vec4_mat4(poly[i], V->mvp, c);
        for (int k = 0; k < 4; ++k) t[k] = c[k] / c[3];
        vec4_mat4(t, V->viewport, scr[i]);
    }
    double ab[3], ac[3], pn[3];
    for (int k = 0; k < 3; ++k) { ab[k] = scr[0][k] - scr[1][k]; ac[k] = scr[0][k] - scr[2][k]; }
    cross3(ab, ac, pn);
    const int is_front = pn[2] < 0;
    double na[3] = {-scr[0][0], -scr[0][1], -scr[0][2]};
    const double D = dot_seq(na, pn, 3);
    int32_t box[4];
    if (!bound_box(&scr[0][0], n, H, W, box)) return;
    for (int px = box[0]; px < box[1]; ++px) for (int py = box[2]; py < box[3]; ++py) {
        int in = 1;
        for (int i = 0; i < n && in; ++i) {
            const double* p0 = scr[i]; const double* p1 = scr[(i + 1) % n];
            double cr = ((double)px - p0[0]) * (p1[1] - p0[1]) - ((double)py - p0[1]) * (p1[0] - p0[0]);
            in = is_front ? (cr > 0) : (cr < 0);
        }
        if (!in) continue;
        double z = -(pn[0] * (double)px + pn[1] * (double)py + D) / pn[2];
        z = linearize_z(z, V->near_, V->far_);
        const int64_t pix = (int64_t)py * W + px;
        const double zb = C->z[pix];
        const int pass = V->system == 1 ? (zb >= z) : (zb <= z);
        if (pass) C->stencil[pix] = (int16_t)(C->stencil[pix] + (is_front ? 1 : -1));
    }
}

/* ---- skybox (cube_map.py:63-101) --------------------------------------------------------------------------- */
static void skybox_fill(ctx_t* C) {
    const b2r_view* V = C->view;
    const int H = C->H, W = C->W, S = C->sky->size;
    static const double corner[2][3][4] = {{{-1, 1, 1, 1}, {1, 1, 1, 1}, {-1, -1, 1, 1}},
                                           {{1, 1, 1, 1}, {1, -1, 1, 1}, {-1, -1, 1, 1}}};
    for (int t = 0; t < 2; ++t) {
        int64_t a[3][2]; double rays[3][3];
        for (int i = 0; i < 3; ++i) {
            double s[4], r[4];
            vec4_mat4(corner[t][i], V->viewport, s);
            a[i][0] = (int64_t)s[0]; a[i][1] = (int64_t)s[1]; /* .astype(int) truncation */
            vec4_mat4(corner[t][i], V->sky_inv, r);
            for (int k = 0; k < 3; ++k) rays[i][k] = r[k] / r[3];
        }
        /* barycentric on integer arrays: the dots are exact int64, then rounded to float32 */
        const int64_t v0x = a[1][0] - a[0][0], v0y = a[1][1] - a[0][1], v1x = a[2][0] - a[0][0], v1y = a[2][1] - a[0][1];
        const float d00 = (float)(v0x * v0x + v0y * v0y), d01 = (float)(v0x * v1x + v0y * v1y),
                    d11 = (float)(v1x * v1x + v1y * v1y);
        const float denom = d00 * d11 - d01 * d01;
        if (denom == 0) continue; /* the reference would raise TypeError on `None >= 0` */
        const float inv = 1.0f / denom;
        for (int py = 0; py < H; ++py) for (int px = 0; px < W; ++px) {
            const int64_t v2x = px - a[0][0], v2y = py - a[0][1];
            const float d20 = (float)(v2x * v0x + v2y * v0y), d21 = (float)(v2x * v1x + v2y * v1y);
            const float bv = (d11 * d20 - d01 * d21) * inv, bw = (d00 * d21 - d01 * d20) * inv, bu = 1.0f - bv - bw;
            if (!(bu >= 0 && bv >= 0 && bw >= 0)) continue;
            double b[3] = {(double)bu, (double)bv, (double)bw}, r[3];
            for (int k = 0; k < 3; ++k) r[k] = seq3(b[0], b[1], b[2], rays[0][k], rays[1][k], rays[2][k]);
            /* CubeMap.__getitem__ (cube_map.py:63-80) */
            int axis = 0; double best = fabs(r[0]);
            if (fabs(r[1]) > best) { best = fabs(r[1]); axis = 1; }
            if (fabs(r[2]) > best) { best = fabs(r[2]); axis = 2; }
            const double amp = r[axis];
            const double o0 = r[axis == 0 ? 1 : 0], o1 = r[axis == 2 ? 1 : 2];
            const double u0 = (o0 / amp + 1) / 2, u1 = (o1 / amp + 1) / 2;
            const int side = (amp < 0) + axis * 2;
            int64_t i0 = (int64_t)(u0 * S - 1), i1 = (int64_t)(u1 * S - 1);
            if (i0 < 0) i0 += S; if (i1 < 0) i1 += S;
            if (i0 < 0 || i0 >= S) i0 = 0; if (i1 < 0 || i1 >= S) i1 = 0;
            const uint8_t* tx = C->sky->faces + (((int64_t)side * S + i0) * S + i1) * 3;
            float* f = C->frame + ((int64_t)py * W + px) * 3;
            for (int k = 0; k < 3; ++k) f[k] = (float)((double)tx[k] / 255);
        }
    }
}

/* ---------------------------------------------------------------------------------------------------------- */
/* One frame.  out_rgb (H,W,3) final image rows; optional z/stencil/winner planes in BUFFER row order;
 * face_status: total_faces bytes (pass-3 status); n_sil: n_models ints.
 * sil_state: optional persistent silhouette, an array of (a,b) pairs per model (in/out) -- NULL = fresh set. */
int orc_render_view(const b2r_model_desc* models, int32_t n_models, const b2r_texture_desc* textures,
                    int32_t n_textures, const b2r_cubemap_desc* sky, const b2r_frame_params* fp,
                    const b2r_view* view, uint8_t* out_rgb, double* out_z, int16_t* out_stencil,
                    int32_t* out_winner, uint8_t* face_status, int32_t* n_sil, float* out_frame_f32,
                    int32_t* out_winner1) {
    ctx_t C;
    memset(&C, 0, sizeof(C));
    C.models = models; C.n_models = n_models; C.textures = textures; C.n_textures = n_textures;
    C.sky = sky; C.fp = fp; C.view = view; C.H = fp->height; C.W = fp->width;
    const int H = C.H, W = C.W;
    const int64_t npx = (int64_t)H * W;
    for (int i = 0; i < 256; ++i) {
        double t = (double)i / 255;
        C.lut_unorm[i] = (float)t;
        C.lut_snorm[i] = (float)(t * 2 - 1);
    }
    C.frame = (float*)calloc((size_t)npx * 3, sizeof(float));
    C.z = (double*)malloc((size_t)npx * sizeof(double));
    C.stencil = (int16_t*)calloc((size_t)npx, sizeof(int16_t));
    C.winner1 = (int32_t*)malloc((size_t)npx * sizeof(int32_t));
    C.winner3 = (int32_t*)malloc((size_t)npx * sizeof(int32_t));
    for (int64_t i = 0; i < npx; ++i) { C.z[i] = view->system == 1 ? INFINITY : -INFINITY; C.winner1[i] = C.winner3[i] = -1; }
    if (fp->bg_mode == B2R_BG_CUBEMAP && sky) skybox_fill(&C);
    else for (int64_t i = 0; i < npx; ++i) for (int k = 0; k < 3; ++k) C.frame[i * 3 + k] = fp->background[k];

    edgeset_t* sets = (edgeset_t*)calloc((size_t)n_models, sizeof(edgeset_t));
    /* pass 1 (core.py:603-606) */
    int face_base = 0;
    for (int mi = 0; mi < n_models; ++mi) {
        const b2r_model_desc* m = models + mi;
        int cap = 64; while (cap < m->n_faces * 8) cap <<= 1;
        edgeset_init(&sets[mi], cap);
        for (int f = 0; f < m->n_faces; ++f) {
            face_t F; face_load(m, f, &F);
            double n[3]; unit_normal_world(&F, n);
            if (dot_seq(n, fp->light.position, 3) > 0)
                for (int i = 0; i < 3; ++i) edgeset_toggle(&sets[mi], F.vi[i], F.vi[(i + 1) % 3]);
            rasterize(&C, &F, face_base + f, 0);
        }
        face_base += m->n_faces;
    }
    /* pass 2 (core.py:610-622) */
    const b2r_light* L = &fp->light;
    for (int mi = 0; mi < n_models; ++mi) {
        const b2r_model_desc* m = models + mi;
        if (n_sil) n_sil[mi] = sets[mi].count;
        for (int s = 0; s < sets[mi].cap; ++s) {
            if (sets[mi].used[s] != 1) continue;
            double A[4], B[4], Cc[4], Dd[4];
            int32_t ia = sets[mi].e[s].a, ib = sets[mi].e[s].b;
            if (ia < 0) ia += m->n_vertices; if (ib < 0) ib += m->n_vertices;
            for (int k = 0; k < 4; ++k) {
                A[k] = load_real(m->vertices, m->vertex_dtype, (int64_t)ia * 4 + k);
                B[k] = load_real(m->vertices, m->vertex_dtype, (int64_t)ib * 4 + k);
            }
            if (L->type == B2R_LIGHT_POINT) {
                const double lp[4] = {L->position[0], L->position[1], L->position[2], 1};
                const double* src[2] = {A, B}; double* dst[2] = {Cc, Dd};
                for (int q = 0; q < 2; ++q) {
                    double dv[4]; for (int k = 0; k < 4; ++k) dv[k] = src[q][k] - lp[k];
                    double l = sqrt(((dv[0] * dv[0] + dv[1] * dv[1]) + dv[2] * dv[2]) + dv[3] * dv[3]);
                    if (l == 0) l = 1;
                    for (int k = 0; k < 4; ++k) dst[q][k] = src[q][k] + 1000 * (dv[k] / l);
                }
            } else {
                const double off[4] = {L->direction[0] * -1000, L->direction[1] * -1000, L->direction[2] * -1000, 1};
                for (int k = 0; k < 4; ++k) { Cc[k] = A[k] + off[k]; Dd[k] = B[k] + off[k]; }
            }
            double quad[16];
            memcpy(quad, A, 32); memcpy(quad + 4, B, 32); memcpy(quad + 8, Dd, 32); memcpy(quad + 12, Cc, 32);
            raster_quad(&C, quad);
        }
    }
    /* pass 3 (core.py:624-636) */
    face_base = 0;
    for (int mi = 0; mi < n_models; ++mi) {
        const b2r_model_desc* m = models + mi;
        for (int f = 0; f < m->n_faces; ++f) {
            face_t F; face_load(m, f, &F);
            int st = rasterize(&C, &F, face_base + f, 1);
            if (face_status) face_status[face_base + f] = (uint8_t)st;
        }
        face_base += m->n_faces;
        edgeset_free(&sets[mi]);
    }
    free(sets);
    /* tonemap + flip (core.py:640) */
    for (int r = 0; r < H; ++r) for (int c = 0; c < W; ++c) for (int k = 0; k < 3; ++k) {
        float v = powf(C.frame[((int64_t)(H - 1 - r) * W + c) * 3 + k], 0.8f) * 255.0f;
        out_rgb[((int64_t)r * W + c) * 3 + k] = (uint8_t)(int32_t)v;
    }
    if (out_z) memcpy(out_z, C.z, (size_t)npx * sizeof(double));
    if (out_stencil) memcpy(out_stencil, C.stencil, (size_t)npx * sizeof(int16_t));
    if (out_winner) for (int64_t i = 0; i < npx; ++i) out_winner[i] = C.winner3[i] >= 0 ? C.winner3[i] : C.winner1[i];
    if (out_winner1) memcpy(out_winner1, C.winner1, (size_t)npx * sizeof(int32_t));
    if (out_frame_f32) memcpy(out_frame_f32, C.frame, (size_t)npx * 3 * sizeof(float));
    free(C.frame); free(C.z); free(C.stencil); free(C.winner1); free(C.winner3);
    return C.index_error ? 2 : 0;  /* 2: the reference would have raised IndexError */
}

/* n_views frames, views rendered concurrently on `threads` host threads (frames are independent: this is how
 * the single-threaded reference scales over cores -- P processes, P different frames).  Debug planes optional. */
typedef struct {
    const b2r_model_desc* models; int32_t n_models; const b2r_texture_desc* textures; int32_t n_textures;
    const b2r_cubemap_desc* sky; const b2r_frame_params* fp; const b2r_view* views; int32_t n_views;
    uint8_t* out_rgb; double* out_z; int16_t* out_stencil; int32_t* out_winner; uint8_t* face_status; int32_t* n_sil;
    int total_faces; int next; int rc; pthread_mutex_t lock;
} job_t;

static void* worker(void* arg) {
    job_t* J = (job_t*)arg;
    const int64_t npx = (int64_t)J->fp->height * J->fp->width;
    for (;;) {
        pthread_mutex_lock(&J->lock);
        int v = J->next++;
        pthread_mutex_unlock(&J->lock);
        if (v >= J->n_views) break;
        const int rc = orc_render_view(J->models, J->n_models, J->textures, J->n_textures, J->sky, J->fp, J->views + v,
                        J->out_rgb + (int64_t)v * npx * 3, J->out_z ? J->out_z + (int64_t)v * npx : 0,
                        J->out_stencil ? J->out_stencil + (int64_t)v * npx : 0,
                        J->out_winner ? J->out_winner + (int64_t)v * npx : 0,
                        J->face_status ? J->face_status + (int64_t)v * J->total_faces : 0,
                        J->n_sil ? J->n_sil + (int64_t)v * J->n_models : 0, 0, 0);
        if (rc) { pthread_mutex_lock(&J->lock); J->rc = rc; pthread_mutex_unlock(&J->lock); }
    }
    return 0;
}

int orc_render(const b2r_model_desc* models, int32_t n_models, const b2r_texture_desc* textures, int32_t n_textures,
               const b2r_cubemap_desc* sky, const b2r_frame_params* fp, const b2r_view* views, int32_t n_views,
               uint8_t* out_rgb, double* out_z, int16_t* out_stencil, int32_t* out_winner, uint8_t* face_status,
               int32_t* n_sil, int32_t threads) {
    job_t J;
    memset(&J, 0, sizeof(J));
    J.models = models; J.n_models = n_models; J.textures = textures; J.n_textures = n_textures; J.sky = sky;
    J.fp = fp; J.views = views; J.n_views = n_views; J.out_rgb = out_rgb; J.out_z = out_z;
    J.out_stencil = out_stencil; J.out_winner = out_winner; J.face_status = face_status; J.n_sil = n_sil;
    for (int mi = 0; mi < n_models; ++mi) J.total_faces += models[mi].n_faces;
    pthread_mutex_init(&J.lock, 0);
    if (threads > n_views) threads = n_views;
    if (threads <= 1) { worker(&J); return J.rc; }
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)threads);
    for (int t = 0; t < threads; ++t) pthread_create(&th[t], 0, worker, &J);
    for (int t = 0; t < threads; ++t) pthread_join(th[t], 0);
    free(th);
    pthread_mutex_destroy(&J.lock);
    return J.rc;
}
