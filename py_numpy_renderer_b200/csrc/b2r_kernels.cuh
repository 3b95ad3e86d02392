// b2r_kernels.cuh -- the sm_100a kernels of the frame pipeline (one translation unit with b2r_api.cu).
//
// Stage map (reference call sites in parentheses):
//   k_facing        per face         light-facing flag, arithmetic in the vertex dtype      (triangular.py:294-295)
//   k_silhouette    per edge         replay of the set toggles in face order, extrusion to a world quad
//                                                                          (triangular.py:296-302, core.py:610-621)
//   k_tri_setup     per face x view  vertex transform, cull, box, f32 barycentric constants, N==1 flags
//                                                                          (triangular.py:36-78)
//   k_quad_setup    per quad x view  Sutherland-Hodgman clip, projection, plane, box
//                                                                          (plane_intersection.py:59-86, triangular.py:319-340)
//   k_bin<FILL>     per primitive    tile lists (count / fill): thread per face, warp per quad with the exact
//                                    corner classification of every (quad, tile) pair
//   k_scan, k_order per view         exclusive scan of the tile counts; tiles bucketed by cost, active tiles counted
//   k_clip_elide    per (screen-filling face, tile)  proof that no pixel of the tile fails the per-pixel clip test: the
//                                    list entry loses its clip bit                           (triangular.py:80-87)
//   k_tile<FUSED>   per 32x32 tile   (1) depth: staged triangle lists, dense (triangle,pixel) dealing, 64-bit keyed
//                                    smem atomics, race stamps  (2) winner: last improver, verified where two
//                                    improvements raced, full pass only on ties  (3) stencil: depth-range classification
//                                    of (quad,tile) pairs, exact row spans, row-level depth classification -> per-row
//                                    difference arrays, the rest as dense per-pixel items
//                                    (4) one packed word per pixel (winner | lit << 31) -- or, FUSED, the shading itself
//                                                                          (triangular.py:78-118, 341-368)
//   k_shade_packed<MODE> per 32x32 tile   Phong + textures + tangent normal maps + skybox + tonemap, warp-packed stores;
//                                    SHADE_F32 (production): float64 coverage / perspective weights / texel addresses,
//                                    float32 lighting; SHADE_F64: all float64; SHADE_ALT: flat / gouraud / pbr;
//                                    tiles without primitives are recognised from their empty lists
//                                                                          (triangular.py:135-171, core.py:138-228,640; cube_map.py:63-101)
//   k_window_push   per 32x32 tile   multi-GPU: sparse push of finished frames into the assembling rank's window (peer stores)
#pragma once
#include "b2r_device.cuh"

// tuning switches of the tile kernel (A/B builds: tools/build_variant.sh <name> -DB2R_...=0)
#ifndef B2R_SKIP
#define B2R_SKIP 0         // timing experiments only (wrong frames): 1 depth pass, 2 stencil phase, 4 winner verification,
#endif                     // 8 span search, 16 everything after the span search of a pair (profiles/r02_phase_timing.md)
#ifndef B2R_CLASSIFY32
#define B2R_CLASSIFY32 0   // 1: one lane per (quad, tile) pair classifies it, up to 32 pairs per grab (measured: diablo 1.90 against
#endif                     // 1.84 ms, torus 1.93 against 2.05 ms per 16 views); 0 (production): four lanes per pair, 8 pairs per grab
#ifndef B2R_ROWDIFF
#define B2R_ROWDIFF 1      // stencil: row-level depth classification + per-row difference arrays (two atomics per row span)
#endif

namespace b2r {

// The reference's float32 texel from the uint8 source (core.py:96-104): f32(u8/255) or f32(u8/255*2-1), float64
// intermediates.  Closed forms that are bit-identical for all 256 inputs (checked exhaustively in
// tests/test_host_api.py::test_texel_decode_closed_forms): a per-lane table lookup in constant memory serialises on
// every distinct index of the warp, these are four instructions.
__device__ __forceinline__ float texel_decode(unsigned u8, int snorm) {
    const double u = (double)u8;
    return snorm ? (float)(u * (2.0 / 255.0) - 1.0) : (float)(u * (1.0 / 255.0));
}

// optional work counters (build with -DB2R_STATS; read with b2r_debug_stats) -- never in the production library
__device__ unsigned long long g_stats[16];
#ifdef B2R_STATS
#define B2R_STAT(i, n) atomicAdd(&g_stats[i], (unsigned long long)(n))
#else
#define B2R_STAT(i, n) ((void)0)
#endif

struct SceneDev {
    const double4* pos;      // (Vtot) world positions, exact promotion of the model's storage
    const double2* uv;       // (Ttot) u, v
    const double* nrm;       // (Ntot,3)
    const FaceStatic* faces; // (F)
    const int4* face_vf;     // (F) v0, v1, v2, FS_* flags: what the per-view set-up needs of a face, 16 B
    const ShadeStatic* shade; // (F) gathered per-face shading inputs
    const ShadeLite* shade_lite; // (F) the same for the float32 lighting path
    const MaterialDev* mats;
    const TextureDev* tex;
    const uchar4* sky;       // (6,S,S) RGBX
    const int2* edge_v;      // (E) canonical (lo, hi) GLOBAL vertex indices
    const int* edge_ptr;     // (E+1)
    const int* edge_inc;     // incidences: face << 1 | reversed, ordered by (face, corner)
    int n_faces, n_edges;
};

// =====================================================================================================================
// light-dependent, view-independent stages
// =====================================================================================================================

// Face.unit_normal_world_space @ light.position > 0   (core.py:127-130, triangular.py:295)
__global__ void k_facing(SceneDev S, LightDev L, uint8_t* __restrict__ facing) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= S.n_faces) return;
    const FaceStatic fs = S.faces[f];
    const double4 a = S.pos[fs.v[0]], b = S.pos[fs.v[1]], c = S.pos[fs.v[2]];
    double n[3];
    if (fs.flags & FS_VTX_F32) {  // cross / norm / divide in float32, every op rounded
        const float e0x = __fsub_rn((float)b.x, (float)a.x), e0y = __fsub_rn((float)b.y, (float)a.y),
                    e0z = __fsub_rn((float)b.z, (float)a.z);
        const float e1x = __fsub_rn((float)c.x, (float)a.x), e1y = __fsub_rn((float)c.y, (float)a.y),
                    e1z = __fsub_rn((float)c.z, (float)a.z);
        const float cx = __fsub_rn(__fmul_rn(e0y, e1z), __fmul_rn(e0z, e1y));
        const float cy = __fsub_rn(__fmul_rn(e0z, e1x), __fmul_rn(e0x, e1z));
        const float cz = __fsub_rn(__fmul_rn(e0x, e1y), __fmul_rn(e0y, e1x));
        float l = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(cx, cx), __fmul_rn(cy, cy)), __fmul_rn(cz, cz)));
        if (l == 0.0f) l = 1.0f;
        n[0] = (double)__fdiv_rn(cx, l); n[1] = (double)__fdiv_rn(cy, l); n[2] = (double)__fdiv_rn(cz, l);
    } else {
        const double e0x = b.x - a.x, e0y = b.y - a.y, e0z = b.z - a.z, e1x = c.x - a.x, e1y = c.y - a.y, e1z = c.z - a.z;
        n[0] = e0y * e1z - e0z * e1y; n[1] = e0z * e1x - e0x * e1z; n[2] = e0x * e1y - e0y * e1x;
        normalize3(n);
    }
    facing[f] = seq3(n[0], n[1], n[2], L.position[0], L.position[1], L.position[2]) > 0 ? 1 : 0;
}

// One thread per undirected edge: replay the set toggles of shadow_volumes() in face order, then extrude.
// state: 0 absent, 1 present as (lo,hi), 2 present as (hi,lo); `persist` != nullptr carries it across renders.
__global__ void k_silhouette(SceneDev S, LightDev L, const uint8_t* __restrict__ facing, int8_t* persist, int toggle,
                             SilEdge* __restrict__ out, int out_cap, int* __restrict__ out_count,
                             int* __restrict__ per_model_count, const int* __restrict__ edge_model) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= S.n_edges) return;
    int state = persist ? persist[e] : 0;
    if (toggle || !persist) {   // toggle == 0: a retry after a capacity overflow, the persistent state is already final
        for (int i = S.edge_ptr[e]; i < S.edge_ptr[e + 1]; ++i) {
            const int inc = S.edge_inc[i];
            if (facing[inc >> 1]) state = state ? 0 : 1 + (inc & 1);
        }
        if (persist) persist[e] = (int8_t)state;
    }
    if (!state) return;
    const int2 ev = S.edge_v[e];
    const double4 A4 = S.pos[state == 1 ? ev.x : ev.y], B4 = S.pos[state == 1 ? ev.y : ev.x];
    const double A[4] = {A4.x, A4.y, A4.z, A4.w}, B[4] = {B4.x, B4.y, B4.z, B4.w};
    double Cc[4], Dd[4];
    if (L.type == B2R_LIGHT_POINT) {
        const double lp[4] = {L.position[0], L.position[1], L.position[2], 1.0};
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const double* s = q ? B : A;
            double* d = q ? Dd : Cc;
            double dv[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) dv[k] = s[k] - lp[k];
            double l = sqrt(((dv[0] * dv[0] + dv[1] * dv[1]) + dv[2] * dv[2]) + dv[3] * dv[3]);
            if (l == 0) l = 1;
#pragma unroll
            for (int k = 0; k < 4; ++k) d[k] = s[k] + 1000.0 * (dv[k] / l);
        }
    } else {  // DIRECTIONAL and SPOT: + (-1000*direction, 1)  => w = 2  (core.py:617-619)
        const double off[4] = {L.direction[0] * -1000.0, L.direction[1] * -1000.0, L.direction[2] * -1000.0, 1.0};
#pragma unroll
        for (int k = 0; k < 4; ++k) { Cc[k] = A[k] + off[k]; Dd[k] = B[k] + off[k]; }
    }
    const int slot = atomicAdd(out_count, 1);
    atomicAdd(per_model_count + edge_model[e], 1);
    if (slot >= out_cap) return;   // capacity overflow: the count keeps growing, the host raises the capacity and retries
    SilEdge& o = out[slot];
#pragma unroll
    for (int k = 0; k < 4; ++k) { o.q[k] = A[k]; o.q[4 + k] = B[k]; o.q[8 + k] = Dd[k]; o.q[12 + k] = Cc[k]; }
}

// =====================================================================================================================
// per-view primitive setup
// =====================================================================================================================
__device__ __forceinline__ void load_clip_coords(const SceneDev& S, const ViewDev& V, const FaceStatic& fs, double* cc) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double4 p = S.pos[fs.v[i]];
        const double w[4] = {p.x, p.y, p.z, p.w};
        vec4_mat4(w, V.mvp, cc + i * 4);
        vec4_mat4(w, V.mvp_dbg, cc + 12 + i * 4);
    }
}

// covered & unclipped test of one pixel, the predicate `Bi` of triangular.py:78-87
__device__ __forceinline__ bool tri_pixel_in(const TriRec& r, const double* cc, bool need_clip, int px, int py,
                                             float& bu, float& bv, float& bw) {
    bool in = tri_bary(r, px, py, bu, bv, bw);
    if (in && need_clip) {
        double P[3];
        persp_bary(r, bu, bv, bw, (r.flags & TR_BOX_ONE) != 0, P);
        in = pixel_unclipped(cc, P, (r.flags & TR_BOX_ONE) != 0);
    }
    return in;
}
__device__ __forceinline__ bool tri_pixel_in(const TriRec& r, const double* cc, int px, int py, float& bu, float& bv,
                                             float& bw) {
    return tri_pixel_in(r, cc, (r.flags & TR_NEEDS_CLIP) != 0, px, py, bu, bv, bw);
}

constexpr int COV_SERIAL_MAX = 128;  // boxes up to this many pixels are counted by the owning thread

// One thread per (vertex, view): the vertex stage of rasterize() (triangular.py:36-45) evaluated once per vertex
// instead of once per corner of every incident face (a vertex of a closed mesh has ~6 of them).  The operations and
// their order are exactly those of the per-face form, so the records are the same bits.
//   rec = (screen x, screen y, screen z (viewport), 1/w)      inside = all of |x|,|y|,|z| < w(1-1e-9) in BOTH frusta
__global__ void k_vertex(SceneDev S, int n_vertices, const ViewDev* __restrict__ views, double4* __restrict__ vrec,
                         uint8_t* __restrict__ vinside, int view0) {
    const int view = blockIdx.y + view0;
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_vertices) return;
    const ViewDev& V = views[view];
    const double4 p = S.pos[v];
    const double w[4] = {p.x, p.y, p.z, p.w};
    double c[4], cd[4];
    vec4_mat4(w, V.mvp, c);
    vec4_mat4(w, V.mvp_dbg, cd);
    const double m = 1.0 - 1e-9;
    const double wc = c[3] * m, wd = cd[3] * m;
    const bool inside = wc > 0 && wd > 0 && fabs(c[0]) < wc && fabs(c[1]) < wc && fabs(c[2]) < wc &&
                        fabs(cd[0]) < wd && fabs(cd[1]) < wd && fabs(cd[2]) < wd;
    const double dpt = 1.0 / c[3];
    const double t[4] = {c[0] * dpt, c[1] * dpt, c[2] * dpt, c[3] * dpt};
    double s4[4];
    vec4_mat4(t, V.viewport, s4);
    vrec[(size_t)view * n_vertices + v] = make_double4(s4[0], s4[1], s4[2], dpt);
    vinside[(size_t)view * n_vertices + v] = inside ? 1 : 0;
}

// One thread per (face, view).  The common path keeps no clip coordinates alive (the per-pixel clip test is elided for
// faces well inside both frusta): 3 x (2 vec.mat, 1/w, viewport) -> cull -> box -> float32
// constants -> N == 1 count over the (small) box.  Faces that need the per-pixel clip test for the count, or whose
// box is large, are queued for k_tri_count (a warp per face).  Every face writes its 16-byte TriBox; only valid faces
// touch their 128-byte record.
__global__ void __launch_bounds__(128, 8)
k_tri_setup(SceneDev S, int n_vertices, const ViewDev* __restrict__ views, FrameDev Fr,
            const double4* __restrict__ vrec, const uint8_t* __restrict__ vinside, TriRec* __restrict__ recs,
            TriBox* __restrict__ boxes, uint8_t* __restrict__ status, int* __restrict__ coop_count,
            int* __restrict__ coop_list, int view0) {
    const int view = blockIdx.y + view0;
    const ViewDev& V = views[view];
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= S.n_faces) return;
    const int4 fv = S.face_vf[f];   // v0, v1, v2, FS_* flags
    const int vidx[3] = {fv.x, fv.y, fv.z};
    TriRec r;
    double sx[3], sy[3], sz[3];
    bool inside = true;   // all three vertices well inside both frusta: the per-pixel clip test cannot fail
#pragma unroll
    for (int i = 0; i < 3; ++i) {  // the vertex stage (triangular.py:36-45) was evaluated per vertex by k_vertex
        const double4 q = vrec[(size_t)view * n_vertices + vidx[i]];
        sx[i] = q.x; sy[i] = q.y; sz[i] = q.z;
        r.d[i] = q.w;
        inside = inside && vinside[(size_t)view * n_vertices + vidx[i]] != 0;
    }
    int st = -1;
    bool need_coop = false;
    if (V.backface) {  // normalize(cross(b-a, c-a))[2] < 0 on screen xyz (triangular.py:47, core.py:132-136)
        const double e0x = sx[1] - sx[0], e0y = sy[1] - sy[0], e0z = sz[1] - sz[0];
        const double e1x = sx[2] - sx[0], e1y = sy[2] - sy[0], e1z = sz[2] - sz[0];
        double n[3] = {e0y * e1z - e0z * e1y, e0z * e1x - e0x * e1z, e0x * e1y - e0y * e1x};
        normalize3(n);
        if (n[2] < 0) st = B2R_FACE_BACK_FACE_CULLING;
    }
    if (st < 0) {  // bound_box (transformation.py:35-43)
        double mnx = fmin(fmin(sx[0], sx[1]), sx[2]), mxx = fmax(fmax(sx[0], sx[1]), sx[2]);
        double mny = fmin(fmin(sy[0], sy[1]), sy[2]), mxy = fmax(fmax(sy[0], sy[1]), sy[2]);
        mnx = mnx < 0 ? 0 : mnx; mxx = mxx > Fr.W ? (double)Fr.W : mxx;
        mny = mny < 0 ? 0 : mny; mxy = mxy > Fr.H ? (double)Fr.H : mxy;
        if (mnx > mxx || mny > mxy || !(mnx == mnx) || !(mxx == mxx) || !(mny == mny) || !(mxy == mxy)) {
            st = B2R_FACE_EMPTY_Z;
        } else {
            r.bx0 = (short)(int)ceil(mnx); r.bx1 = (short)(int)ceil(mxx);
            r.by0 = (short)(int)ceil(mny); r.by1 = (short)(int)ceil(mxy);
        }
    }
    if (st < 0) {  // barycentric constants (transformation.py:16-28)
        r.ax = sx[0]; r.ay = sy[0];
        r.v0x = sx[1] - sx[0]; r.v0y = sy[1] - sy[0];
        r.v1x = sx[2] - sx[0]; r.v1y = sy[2] - sy[0];
        r.d00 = (float)seq2(r.v0x, r.v0y, r.v0x, r.v0y);
        r.d01 = (float)seq2(r.v0x, r.v0y, r.v1x, r.v1y);
        r.d11 = (float)seq2(r.v1x, r.v1y, r.v1x, r.v1y);
        const float den = __fsub_rn(__fmul_rn(r.d00, r.d11), __fmul_rn(r.d01, r.d01));
        if (den == 0.0f) st = B2R_FACE_EMPTY_B;
        else r.inv = __fdiv_rn(1.0f, den);
    }
    if (st < 0) {
        const int nx = r.bx1 > r.bx0 ? r.bx1 - r.bx0 : 0, ny = r.by1 > r.by0 ? r.by1 - r.by0 : 0;
        const int n_box = nx * ny;
        if (n_box == 0) st = B2R_FACE_CLIPPED;
        else {
            r.flags = TR_VALID | (n_box == 1 ? TR_BOX_ONE : 0) | ((fv.w & FS_NO_ZWRITE) ? TR_NO_ZWRITE : 0) |
                      (((fv.w & FS_CLIP) && !inside) ? TR_NEEDS_CLIP : 0);
#pragma unroll
            for (int i = 0; i < 3; ++i) r.zl[i] = linearize_z(sz[i], V);
            r.pad = 0;
            // N of `bar_screen[Bi]` decides the evaluation order of the z interpolation: count covered & unclipped
            // pixels, stopping at 2.  Small boxes without a clip test are counted right here.
            if (n_box <= COV_SERIAL_MAX && !(r.flags & TR_NEEDS_CLIP)) {
                int cnt = 0, px = r.bx0, py = r.by0;
                for (int i = 0; i < n_box && cnt < 2; ++i) {
                    float bu, bv, bw;
                    cnt += tri_bary(r, px, py, bu, bv, bw) ? 1 : 0;
                    if (++py == r.by1) { py = r.by0; ++px; }
                }
                if (cnt == 1) r.flags |= TR_COV_ONE;
                if (cnt == 0) { st = B2R_FACE_CLIPPED; r.flags = 0; }
            } else {
                need_coop = true;
            }
            if (st < 0) recs[(size_t)view * S.n_faces + f] = r;
        }
    }
    TriBox bx;
    bx.pad = 0;
    if (st >= 0) {   // culled / empty / clipped: the 128-byte record is not touched at all
        bx.bx0 = bx.bx1 = bx.by0 = bx.by1 = 0; bx.flags = 0;
        if (status) status[(size_t)view * S.n_faces + f] = (uint8_t)st;
    } else {
        bx.bx0 = r.bx0; bx.bx1 = r.bx1; bx.by0 = r.by0; bx.by1 = r.by1; bx.flags = r.flags;
        if (status) status[(size_t)view * S.n_faces + f] = 0xE1;  // pending + covered: the tile kernel ORs the z / lit bits in
    }
    boxes[(size_t)view * S.n_faces + f] = bx;
    if (need_coop) coop_list[(size_t)view * S.n_faces + atomicAdd(coop_count + view, 1)] = f;
}

// N == 1 count of the faces k_tri_setup queued (large boxes, faces that need the per-pixel clip test): a warp per face,
// lanes stride over the box in a scattered order and stop at two hits.
__global__ void k_tri_count(SceneDev S, const ViewDev* __restrict__ views, TriRec* __restrict__ recs,
                            TriBox* __restrict__ boxes, uint8_t* __restrict__ status, const int* __restrict__ coop_count,
                            const int* __restrict__ coop_list, int view0) {
    const int view = blockIdx.y + view0;
    const ViewDev& V = views[view];
    const int lane = threadIdx.x & 31;
    const int n_queued = coop_count[view];
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int qi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; qi < n_queued; qi += warps) {
        const int sf = coop_list[(size_t)view * S.n_faces + qi];
        TriRec* my = recs + (size_t)view * S.n_faces + sf;
        const TriRec r = *my;
        double c2[CLIP_DOUBLES];
        if (r.flags & TR_NEEDS_CLIP) load_clip_coords(S, V, S.faces[sf], c2);
        const int ny = r.by1 - r.by0, n_box = (r.bx1 - r.bx0) * ny;
        // visit the box in a scattered order (i * prime mod n is a permutation): a triangle that covers a few
        // percent of its box yields two hits within the first iterations, the scan stays exhaustive otherwise
        const long long stride = (n_box % 1000003) ? 1000003 : 999983;
        int j = (int)(((long long)lane * stride) % n_box);          // pixel i = base + lane sits at (i * stride) % n_box:
        const int step = (int)((32 * stride) % n_box);              // advanced incrementally, no division in the loop
        const double inv_ny = 1.0 / (double)ny;
        int cnt = 0;
        for (int base = 0; base < n_box && cnt < 2; base += 32) {
            bool in = false;
            if (base + lane < n_box) {
                int q = (int)((double)j * inv_ny), rem = j - q * ny;  // j / ny, j % ny (estimate off by at most one)
                if (rem >= ny) { ++q; rem -= ny; } else if (rem < 0) { --q; rem += ny; }
                float bu, bv, bw;
                in = tri_pixel_in(r, c2, r.bx0 + q, r.by0 + rem, bu, bv, bw);
            }
            j += step;
            if (j >= n_box) j -= n_box;
            cnt += __popc(__ballot_sync(0xffffffffu, in));
        }
        if (lane == 0) {
            if (cnt == 1) { my->flags = r.flags | TR_COV_ONE; boxes[(size_t)view * S.n_faces + sf].flags = r.flags | TR_COV_ONE; }
            if (cnt == 0) {
                my->flags = 0;
                boxes[(size_t)view * S.n_faces + sf].flags = 0;
                if (status) status[(size_t)view * S.n_faces + sf] = B2R_FACE_CLIPPED;
            }
        }
    }
}

// Sutherland-Hodgman against the six camera planes in homogeneous world space (plane_intersection.py:59-86),
// then projection and plane set-up of resterize_quadrangle (triangular.py:325-340).  One thread per quad.
__device__ void quad_setup_one(const SilEdge& sil, const ViewDev& V, const FrameDev& Fr, QuadRec& R);
__global__ void k_quad_setup(const SilEdge* __restrict__ sil, const int* __restrict__ sil_count,
                             const ViewDev* __restrict__ views, FrameDev Fr, QuadRec* __restrict__ recs, int rec_stride,
                             int view0) {
    const int view = blockIdx.y + view0;
    const ViewDev& V = views[view];
    const int n_quads = min(*sil_count, rec_stride);   // rec_stride = capacity of the silhouette / quad records
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n_quads; q += gridDim.x * blockDim.x)
        quad_setup_one(sil[q], V, Fr, recs[(size_t)view * rec_stride + q]);
}

__device__ void quad_setup_one(const SilEdge& sil, const ViewDev& V, const FrameDev& Fr, QuadRec& R) {
    constexpr int CAP = 2 * B2R_MAX_POLY;
    double bufA[CAP][4], bufB[CAP][4];
    double (*cur)[4] = bufA;
    double (*nxt)[4] = bufB;
    int n = 4;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k) cur[i][k] = sil.q[i * 4 + k];
    for (int p = 0; p < 6 && n > 0; ++p) {
        const double* pl = V.planes + p * 4;
        int m = 0;
        for (int i = 0; i < n; ++i) {
            const double* c = cur[i];
            const double* nx = cur[(i + 1 == n) ? 0 : i + 1];
            const bool cv = dot4_seq(pl, c) >= 0, nv = dot4_seq(pl, nx) >= 0;
            if (cv && m < CAP) { for (int k = 0; k < 4; ++k) nxt[m][k] = c[k]; ++m; }
            if (cv != nv) {  // line_plane_intersection(next, current, plane)  (plane_intersection.py:24-36)
                double dir[4];
                for (int k = 0; k < 4; ++k) dir[k] = c[k] - nx[k];
                const double den = dot4_seq(pl, dir);
                if (!(fabs(den) < 1e-10)) {
                    const double wgt = -dot4_seq(pl, nx) / den;
                    if (0 <= wgt && wgt <= 1 && m < CAP) {
                        for (int k = 0; k < 4; ++k) nxt[m][k] = nx[k] + wgt * dir[k];
                        ++m;
                    }
                }
            }
        }
        n = m;
        double (*t)[4] = cur; cur = nxt; nxt = t;
    }
    if (n < 3 || n > B2R_MAX_POLY) { R.n = 0; return; }
    double sx[B2R_MAX_POLY], sy[B2R_MAX_POLY], sz[3];
    for (int i = 0; i < n; ++i) {
        double c[4], t[4], s[4];
        vec4_mat4(cur[i], V.mvp, c);
        for (int k = 0; k < 4; ++k) t[k] = c[k] / c[3];
        vec4_mat4(t, V.viewport, s);
        sx[i] = s[0]; sy[i] = s[1];
        if (i < 3) sz[i] = s[2];
    }
    const double abx = sx[0] - sx[1], aby = sy[0] - sy[1], abz = sz[0] - sz[1];
    const double acx = sx[0] - sx[2], acy = sy[0] - sy[2], acz = sz[0] - sz[2];
    const double nx_ = aby * acz - abz * acy, ny_ = abz * acx - abx * acz, nz_ = abx * acy - aby * acx;
    R.front = nz_ < 0 ? 1 : 0;
    R.nx = nx_; R.ny = ny_; R.nz = nz_;
    R.D = seq3(-sx[0], -sy[0], -sz[0], nx_, ny_, nz_);
    double mnx = sx[0], mxx = sx[0], mny = sy[0], mxy = sy[0];
    bool bad = false;
    for (int i = 0; i < n; ++i) {
        R.x[i] = sx[i]; R.y[i] = sy[i];
        bad = bad || !(sx[i] == sx[i]) || !(sy[i] == sy[i]);
        mnx = fmin(mnx, sx[i]); mxx = fmax(mxx, sx[i]); mny = fmin(mny, sy[i]); mxy = fmax(mxy, sy[i]);
    }
    mnx = mnx < 0 ? 0 : mnx; mxx = mxx > Fr.W ? (double)Fr.W : mxx;
    mny = mny < 0 ? 0 : mny; mxy = mxy > Fr.H ? (double)Fr.H : mxy;
    if (bad || mnx > mxx || mny > mxy) { R.n = 0; return; }
    R.bx0 = (short)(int)ceil(mnx); R.bx1 = (short)(int)ceil(mxx);
    R.by0 = (short)(int)ceil(mny); R.by1 = (short)(int)ceil(mxy);
    R.n = (R.bx1 > R.bx0 && R.by1 > R.by0) ? n : 0;
}

// =====================================================================================================================
// binning: per-tile primitive lists (count -> scan -> fill)
// =====================================================================================================================
// Strict inside test of resterize_quadrangle for one polygon edge (triangular.py:305-316), exactly as NumPy
// evaluates np.cross on 2-vectors: (px-x0)*(y1-y0) - (py-y0)*(x1-x0), separate roundings.
__device__ __forceinline__ double edge_fn(double px, double py, double x0, double y0, double ex, double ey) {
    return (px - x0) * ey - (py - y0) * ex;
}

// How does the pixel rectangle [x0,x1] x [y0,y1] (inclusive) meet the polygon?  edge_fn is a composition of monotone
// roundings, so its extremes over the rectangle are attained, EXACTLY, at corners:
//   0 = no pixel of the rectangle can be inside,  1 = some may be,  2 = every pixel of the rectangle is inside.
constexpr int QUAD_FULL_BIT = 1 << 30;  // flag in a tile-list entry: the quad covers every pixel of the tile
constexpr int TRI_CLIP_BIT = 1 << 30;   // flag in a triangle tile-list entry: the face needs the per-pixel clip test
// The polygon of the quad a warp is classifying, staged once in shared memory together with its edge vectors: the
// classification of hundreds of (quad, tile) pairs then reads broadcast shared-memory words instead of the 256-byte
// global record, and the edge subtraction is done once per quad.
struct QuadEdges {
    double x[B2R_MAX_POLY], y[B2R_MAX_POLY], ex[B2R_MAX_POLY], ey[B2R_MAX_POLY];
    int n, front;
};
__device__ __forceinline__ int quad_tile_class(const QuadEdges& R, int x0, int x1, int y0, int y1) {
    bool full = true;
    for (int i = 0; i < R.n; ++i) {
        const double ex = R.ex[i], ey = R.ey[i];
        // corner maximising f = (px-xi)*ey - (py-yi)*ex, and the opposite corner minimising it
        const double fmax = edge_fn(ey >= 0 ? x1 : x0, ex <= 0 ? y1 : y0, R.x[i], R.y[i], ex, ey);
        const double fmin = edge_fn(ey >= 0 ? x0 : x1, ex <= 0 ? y0 : y1, R.x[i], R.y[i], ex, ey);
        if (R.front) {  // inside means f > 0
            if (!(fmax > 0)) return 0;
            full = full && (fmin > 0);
        } else {        // inside means f < 0
            if (!(fmin < 0)) return 0;
            full = full && (fmax < 0);
        }
    }
    return full ? 2 : 1;
}

struct BinDev {
    int* tri_count;   // (views, n_tiles)  count, then running cursor during fill
    int* quad_count;
    int* tri_off;     // (views, n_tiles + 1) exclusive scan
    int* quad_off;
    int* tri_list;    // (views, tri_cap)
    int* quad_list;   // (views, quad_cap)
    int tri_cap, quad_cap;
    int* overflow;    // (views, 2) required sizes when a list does not fit
    int share_cap;    // most warps that split the tiles of one quad
    int2* pair_list;  // (views, quad_cap) surviving (entry, tile) pairs found by the count pass, replayed by the fill pass
    int* pair_count;  // (views)
    int* huge_count;  // (views) faces whose box holds more than BIN_HUGE tiles: the fill pass spreads them over the grid
    int* huge_list;   // (views, BIN_HUGE_CAP)
    int* order;       // (views, n_tiles) tiles sorted by estimated cost class, heaviest first (raster launch order)
    int* n_active;    // (views) how many leading entries of `order` hold tiles the tile kernel has work for
};

// Can any pixel of the rectangle [x0,x1] x [y0,y1] (inclusive) pass the float32 coverage test of `tri_bary`?
// The computed barycentrics are, up to float32 rounding, affine functions of the pixel: b_i = t_i(p) + eps_i with
//   t_v = (d11*D20 - d01*D21)*inv,  t_w = (d00*D21 - d01*D20)*inv,  t_u = 1 - t_v - t_w   (real arithmetic on the
// float32 constants of the record, D2k = (p-a).v_k) and |eps| bounded by the operation-by-operation error analysis
// below (4 roundings at 2^-24 each, generously doubled).  An affine function attains its extremes over a rectangle
// at the corners, so if some t_i stays below -E_i at all four corners, b_i < 0 for every pixel: nothing is covered.
// Only used for big boxes (one screen-filling triangle against the tiles on the far side of its edges).
constexpr int BIG_BOX_PX = 256;
__device__ __forceinline__ bool tri_misses_rect(const TriRec& r, int x0, int x1, int y0, int y1) {
    const double d00 = r.d00, d01 = r.d01, d11 = r.d11, inv = r.inv;
    const double u32 = 5.9604644775390625e-8;  // 2^-24
    double tmax[3] = {-1e300, -1e300, -1e300}, amax_v = 0, amax_w = 0, gv = 0, gw = 0;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const double v2x = (double)((c & 1) ? x1 : x0) - r.ax, v2y = (double)((c & 2) ? y1 : y0) - r.ay;
        const double D20 = v2x * r.v0x + v2y * r.v0y, D21 = v2x * r.v1x + v2y * r.v1y;
        const double tv = (d11 * D20 - d01 * D21) * inv, tw = (d00 * D21 - d01 * D20) * inv, tu = 1.0 - tv - tw;
        if (!(tv == tv) || !(tw == tw)) return false;
        tmax[0] = fmax(tmax[0], tu); tmax[1] = fmax(tmax[1], tv); tmax[2] = fmax(tmax[2], tw);
        amax_v = fmax(amax_v, fabs(tv)); amax_w = fmax(amax_w, fabs(tw));
        gv = fmax(gv, (fabs(d11 * D20) + fabs(d01 * D21)) * fabs(inv));
        gw = fmax(gw, (fabs(d00 * D21) + fabs(d01 * D20)) * fabs(inv));
    }
    const double Ev = 8.0 * u32 * (gv + amax_v) + 1e-30, Ew = 8.0 * u32 * (gw + amax_w) + 1e-30;
    const double Eu = Ev + Ew + 8.0 * u32 * (1.0 + amax_v + amax_w);
    return tmax[0] < -Eu || tmax[1] < -Ev || tmax[2] < -Ew;
}

// Does EVERY pixel of the rectangle [x0,x1] x [y0,y1] (inclusive) that `tri_bary` covers pass the per-pixel clip test of
// triangular.py:80-87 (`pixel_unclipped`) in both frusta?  Then the test -- three float64 divisions and 24 FMAs per
// pixel -- is skipped for this (triangle, tile): the floor of the headline scene reaches far outside the frustum, so
// TR_NEEDS_CLIP is set on its two faces, yet every tile in the interior of the screen passes.
// Why the corners decide: the reference tests -q.w < q.k < q.w (k = x, y, z) on q = P @ clip with P_i = b_i d_i / s,
// s = sum b_j d_j.  With s > 0 this is F = sum_i b_i d_i (clip_i.w -+ clip_i.k) > 0, LINEAR in the barycentrics, and
// the barycentrics are affine in the pixel up to the float32 rounding bounded as in tri_misses_rect (|eps_i| <= E_i).
// An affine function attains its minimum over the rectangle at a corner, so if at all four corners s and the six F of
// each frustum exceed the worst-case perturbation sum_i (E_i + float64 rounding) |d_i| (|clip_i.w| + |clip_i.k|), four
// times over, every computed comparison of every pixel in the rectangle comes out true.  NaN / inf anywhere -> false.
__device__ __forceinline__ bool tri_rect_unclipped(const TriRec& r, const double* __restrict__ cc, int x0, int x1, int y0, int y1) {
    const double d00 = r.d00, d01 = r.d01, d11 = r.d11, inv = r.inv;
    const double u32 = 5.9604644775390625e-8, u64 = 1.1102230246251565e-16;  // 2^-24, 2^-53
    double amax[3] = {0, 0, 0}, gv = 0, gw = 0, smin = 1e300, fmn[2] = {1e300, 1e300};
    bool bad = false;
#pragma unroll
    for (int c = 0; c < 4; ++c) {   // four independent chains: unrolled so that they overlap in the float64 pipe
        const double v2x = (double)((c & 1) ? x1 : x0) - r.ax, v2y = (double)((c & 2) ? y1 : y0) - r.ay;
        const double D20 = v2x * r.v0x + v2y * r.v0y, D21 = v2x * r.v1x + v2y * r.v1y;
        const double tv = (d11 * D20 - d01 * D21) * inv, tw = (d00 * D21 - d01 * D20) * inv, tu = 1.0 - tv - tw;
        bad = bad || !(fabs(tv) < 1e30) || !(fabs(tw) < 1e30);
        amax[0] = fmax(amax[0], fabs(tu)); amax[1] = fmax(amax[1], fabs(tv)); amax[2] = fmax(amax[2], fabs(tw));
        gv = fmax(gv, (fabs(d11 * D20) + fabs(d01 * D21)) * fabs(inv));
        gw = fmax(gw, (fabs(d00 * D21) + fabs(d01 * D20)) * fabs(inv));
        const double e0 = tu * r.d[0], e1 = tv * r.d[1], e2 = tw * r.d[2];
        const double sc = e0 + e1 + e2;
        bad = bad || !(sc > 0);
        smin = fmin(smin, sc);
#pragma unroll
        for (int cam = 0; cam < 2; ++cam) {
            const double* q = cc + cam * 12;
            const double W = e0 * q[3] + e1 * q[7] + e2 * q[11];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const double K = e0 * q[k] + e1 * q[4 + k] + e2 * q[8 + k];
                const double lo = W - K, hi = W + K;
                bad = bad || !(lo > 0) || !(hi > 0);
                fmn[cam] = fmin(fmn[cam], fmin(lo, hi));
            }
        }
    }
    const double Ev = 8.0 * u32 * (gv + amax[1]) + 1e-30, Ew = 8.0 * u32 * (gw + amax[2]) + 1e-30;
    const double E[3] = {Ev + Ew + 8.0 * u32 * (1.0 + amax[1] + amax[2]), Ev, Ew};
    double sb = 0, fb[2] = {0, 0};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double ci = (E[i] + 16.0 * u64 * (amax[i] + E[i])) * fabs(r.d[i]);
        sb += ci;
#pragma unroll
        for (int cam = 0; cam < 2; ++cam) {
            const double* q = cc + cam * 12 + i * 4;
            fb[cam] += ci * (fabs(q[3]) + fmax(fmax(fabs(q[0]), fabs(q[1])), fabs(q[2])));
        }
    }
    return !bad && smin > 4.0 * sb && fmn[0] > 4.0 * fb[0] && fmn[1] > 4.0 * fb[1];
}

// The clip test of a (face, tile) pair decided once, by the thread of the binning fill pass that inserts the pair:
// true when every pixel of the tile that the face can cover passes `pixel_unclipped` (tri_rect_unclipped above), so the
// tile kernel runs that pair without the per-pixel test.  (Deciding it inside the tile kernel was measured first: one
// lane per CTA on a chain of ~600 dependent float64 operations while the other 127 threads wait cost more than the
// test saves.  Here every lane of the fill pass decides its own tile.)
__device__ __forceinline__ bool face_tile_unclipped(const FrameDev& Fr, const TriRec* __restrict__ vtris, const double4* __restrict__ pos,
                                                 const int4* __restrict__ face_vf, const ViewDev& V, int face, int tx, int ty) {
    const TriRec r = vtris[face];
    const int X0 = tx * TILE_W, Y0 = ty * TILE_H;
    const int x0 = max((int)r.bx0, X0), x1 = min((int)r.bx1, min(X0 + TILE_W, Fr.W));
    const int y0 = max((int)r.by0, max(Y0, Fr.row_begin)), y1 = min((int)r.by1, min(min(Y0 + TILE_H, Fr.H), Fr.row_end));
    if ((x1 - x0) * (y1 - y0) < BIG_BOX_PX || x1 <= x0) return false;   // not worth a test: few pixels either way
    const int4 fv = face_vf[face];
    double cc[CLIP_DOUBLES];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double4 p = pos[i == 0 ? fv.x : (i == 1 ? fv.y : fv.z)];
#pragma unroll
        for (int k = 0; k < 4; ++k) {   // the staged form of k_tile (tile_tris): same operations, same order
            cc[i * 4 + k] = fma(p.w, V.mvp[12 + k], fma(p.z, V.mvp[8 + k], fma(p.y, V.mvp[4 + k], p.x * V.mvp[k])));
            cc[12 + i * 4 + k] = fma(p.w, V.mvp_dbg[12 + k], fma(p.z, V.mvp_dbg[8 + k], fma(p.y, V.mvp_dbg[4 + k], p.x * V.mvp_dbg[k])));
        }
    }
    return tri_rect_unclipped(r, cc, x0, x1 - 1, y0, y1 - 1);
}

// One warp per primitive slot: lanes stride over the tiles of the primitive's box.
// Triangles: one THREAD per face; a box touching at most BIN_SMALL tiles is binned by its own thread, bigger ones
// (screen-filling triangles) are handed to the whole warp through a ballot queue, lanes striding over the tiles.
// Shadow quads: one WARP per quad (long slivers crossing many tiles, each with the exact per-tile classification).
constexpr int BIN_SMALL = 4;
constexpr int BIN_HUGE = 256, BIN_HUGE_CAP = 64;
constexpr int BIN_PAIR_BUF = 128;  // per-warp shared-memory run of (quad, tile) pairs between two global appends
constexpr int BIN_SUPER = 4;  // quads: tiles are classified in blocks of BIN_SUPER x BIN_SUPER first
template <bool FILL>
__global__ void k_bin(FrameDev Fr, const TriBox* __restrict__ boxes, const QuadRec* __restrict__ quads,
                      const int* __restrict__ sil_count, int quad_stride, BinDev B, int view0, int parts) {
    // parts: 1 = triangles, 2 = shadow quads, 3 = both.  The count pass runs as two launches when triangle set-up has
    // its own stream: the quad half (the long one) starts right after k_quad_setup, the triangle half after k_tri_count.
    const int view = blockIdx.y + view0;
    const int lane = threadIdx.x & 31;
    const int n_tiles = Fr.tiles_x * Fr.tiles_y;
    const int n_quads = min(*sil_count, quad_stride);   // quad_stride = capacity of the silhouette / quad records
    const int band_y0 = Fr.row_begin, band_y1 = Fr.row_end;
    if (FILL && (B.overflow[view * 2] | B.overflow[view * 2 + 1])) return;
    const int stride = gridDim.x * blockDim.x;
    int* const tri_count = B.tri_count + (size_t)view * n_tiles;
    int* const quad_count = B.quad_count + (size_t)view * n_tiles;
    const int* const tri_off = B.tri_off + (size_t)view * (n_tiles + 1);
    const int* const quad_off = B.quad_off + (size_t)view * (n_tiles + 1);
    int* const tri_list = B.tri_list + (size_t)view * B.tri_cap;
    int* const quad_list = B.quad_list + (size_t)view * B.quad_cap;
    for (int base = blockIdx.x * blockDim.x + threadIdx.x - lane; (parts & 1) && base < Fr.n_faces; base += stride) {
        const int slot = base + lane;
        int bx0 = 0, bx1 = 0, by0 = 0, by1 = 0;
        bool valid = false;
        int entry = slot;   // tile-list entry: face index, TRI_CLIP_BIT set when the face needs the per-pixel clip test
        if (slot < Fr.n_faces) {
            const TriBox r = boxes[(size_t)view * Fr.n_faces + slot];   // one 16-byte load
            if (r.flags & TR_VALID) { valid = true; bx0 = r.bx0; bx1 = r.bx1; by0 = r.by0; by1 = r.by1; }
            if (r.flags & TR_NEEDS_CLIP) entry |= TRI_CLIP_BIT;
        }
        by0 = max(by0, band_y0); by1 = min(by1, band_y1);
        valid = valid && by0 < by1 && bx0 < bx1;
        int tx0 = 0, ty0 = 0, tw = 1, nt = 0;
        if (valid) {
            tx0 = bx0 / TILE_W; ty0 = by0 / TILE_H;
            tw = (bx1 - 1) / TILE_W - tx0 + 1;
            nt = tw * ((by1 - 1) / TILE_H - ty0 + 1);
        }
        const bool small = valid && nt <= BIN_SMALL;
        // Small boxes: the faces of a warp are neighbours in the mesh and mostly fall into the same few tiles, so
        // the lanes that hit the same tile share ONE atomic (match_any): with hundreds of micro-triangles per tile
        // (BASELINE config 5) the per-tile counters are otherwise a same-address serialisation point.
#pragma unroll
        for (int i = 0; i < BIN_SMALL; ++i) {
            const bool act = small && i < nt;
            const unsigned am = __ballot_sync(0xffffffffu, act);
            if (!am) break;
            if (act) {
                const int t = (ty0 + i / tw - Fr.tile_row0) * Fr.tiles_x + tx0 + i % tw;
                const unsigned peers = __match_any_sync(am, t);
                const int leader = __ffs(peers) - 1;
                int base = 0;
                if (lane == leader) base = atomicAdd(tri_count + t, __popc(peers));
                if (FILL) {
                    base = __shfl_sync(peers, base, leader);
                    tri_list[tri_off[t] + base + __popc(peers & ((1u << lane) - 1))] = entry;
                }
            }
        }
        unsigned queue = __ballot_sync(0xffffffffu, valid && !small);
        while (queue) {
            const int src = __ffs(queue) - 1;
            queue &= queue - 1;
            const int s_tx0 = __shfl_sync(0xffffffffu, tx0, src), s_ty0 = __shfl_sync(0xffffffffu, ty0, src);
            const int s_tw = __shfl_sync(0xffffffffu, tw, src), s_nt = __shfl_sync(0xffffffffu, nt, src);
            const int s_entry = __shfl_sync(0xffffffffu, entry, src);
            if (s_nt > BIN_HUGE) {
                // A screen-filling triangle: thousands of list inserts, each waiting for its atomic, would make this
                // one warp the critical path of the fill pass.  The count pass (fire-and-forget atomics) notes the
                // face; the fill pass spreads the noted faces over the whole grid below.
                if (!FILL) {
                    if (lane == 0) {
                        const int h = atomicAdd(B.huge_count + view, 1);
                        if (h < BIN_HUGE_CAP) B.huge_list[view * BIN_HUGE_CAP + h] = base + src;
                    }
                } else {
                    const int nh = min(B.huge_count[view], BIN_HUGE_CAP);
                    const int* hl = B.huge_list + view * BIN_HUGE_CAP;
                    const bool mine = (lane < nh && hl[lane] == base + src) || (lane + 32 < nh && hl[lane + 32] == base + src);
                    if (__any_sync(0xffffffffu, mine)) continue;  // noted: handled by the grid-wide loop
                }
            }
            for (int i = lane; i < s_nt; i += 32) {
                const int t = (s_ty0 + i / s_tw - Fr.tile_row0) * Fr.tiles_x + s_tx0 + i % s_tw;
                if (FILL) tri_list[tri_off[t] + atomicAdd(tri_count + t, 1)] = s_entry;
                else atomicAdd(tri_count + t, 1);
            }
        }
    }
    // A quad whose box spans the screen has thousands of tiles to classify, each a chain of dependent float64
    // operations: the grid's warps are split evenly over the quads (`share` warps each, striding over the tiles of
    // the box) so the launch is not as long as the biggest quad.
    if (FILL) {
        const int nh = min(B.huge_count[view], BIN_HUGE_CAP);
        for (int h = 0; h < nh; ++h) {   // the noted screen-filling faces: one tile per thread of the grid
            const int face = B.huge_list[view * BIN_HUGE_CAP + h];
            const TriBox r = boxes[(size_t)view * Fr.n_faces + face];
            const int h_entry = face | ((r.flags & TR_NEEDS_CLIP) ? TRI_CLIP_BIT : 0);
            const int bx0 = r.bx0, bx1 = r.bx1, by0 = max((int)r.by0, band_y0), by1 = min((int)r.by1, band_y1);
            const int tx0 = bx0 / TILE_W, ty0 = by0 / TILE_H, tw = (bx1 - 1) / TILE_W - tx0 + 1;
            const int nt = tw * ((by1 - 1) / TILE_H - ty0 + 1);
            for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nt; i += stride) {
                const int t = (ty0 + i / tw - Fr.tile_row0) * Fr.tiles_x + tx0 + i % tw;
                tri_list[tri_off[t] + atomicAdd(tri_count + t, 1)] = h_entry;
            }
        }
        // the count pass left the surviving (quad, tile) pairs behind: the fill pass only scatters them
        const int n_pairs = min(B.pair_count[view], B.quad_cap);
        const int2* pairs = B.pair_list + (size_t)view * B.quad_cap;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pairs; i += stride) {
            const int2 pr = pairs[i];
            quad_list[quad_off[pr.y] + atomicAdd(quad_count + pr.y, 1)] = pr.x;
        }
        return;
    }
    if (!(parts & 2)) return;
    // Count pass.  A quad whose box spans the screen has hundreds of tiles to classify, each a chain of dependent
    // float64 operations: the grid's warps are split evenly over the quads (`share` warps each, striding over the
    // blocks of the box).  Surviving pairs collect in a per-warp shared-memory run and reach the per-view pair list
    // with one global atomic per run.
    __shared__ int2 pair_buf[8][BIN_PAIR_BUF];
    __shared__ QuadEdges edges[8];
    int2* const my_buf = pair_buf[threadIdx.x >> 5];
    QuadEdges& QE = edges[threadIdx.x >> 5];
    int2* const pair_list = B.pair_list + (size_t)view * B.quad_cap;
    int n_buf = 0;  // warp-uniform
    auto flush = [&]() {
        int base = 0;
        if (lane == 0) base = atomicAdd(B.pair_count + view, n_buf);
        base = __shfl_sync(0xffffffffu, base, 0);
        __syncwarp();
        for (int i = lane; i < n_buf; i += 32) if (base + i < B.quad_cap) pair_list[base + i] = my_buf[i];
        __syncwarp();
        n_buf = 0;
    };
    const int total_warps = stride >> 5;
    const int share = max(1, min(B.share_cap, total_warps / max(n_quads, 1)));
    const int warp_id = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int part = warp_id % share;
    for (int prim = warp_id / share; prim < n_quads; prim += total_warps / share) {
        const QuadRec* Q = quads + (size_t)view * quad_stride + prim;
        const int qn = Q->n;
        if (qn == 0) continue;
        const int bx0 = Q->bx0, bx1 = Q->bx1, by0 = max((int)Q->by0, band_y0), by1 = min((int)Q->by1, band_y1);
        if (by0 >= by1 || bx0 >= bx1) continue;
        __syncwarp();   // the previous quad's edges are no longer read
        if (lane < qn) { QE.x[lane] = Q->x[lane]; QE.y[lane] = Q->y[lane]; }
        if (lane == 0) { QE.n = qn; QE.front = Q->front; }
        __syncwarp();
        if (lane < qn) {
            const int j = (lane + 1 == qn) ? 0 : lane + 1;
            QE.ex[lane] = QE.x[j] - QE.x[lane]; QE.ey[lane] = QE.y[j] - QE.y[lane];
        }
        __syncwarp();
        // Two levels.  The box of a shadow quad (a long diagonal sliver) holds many more tiles than the quad touches,
        // so blocks of BIN_SUPER x BIN_SUPER tiles are classified first with the same exact corner test: a rejected
        // block rejects its tiles, a block that is inside everywhere makes its tiles "full" without further tests
        // (both follow from the monotonicity argument above, the block's rectangle containing the tiles').
        const int tx0 = bx0 / TILE_W, ty0 = by0 / TILE_H, tx1 = (bx1 - 1) / TILE_W + 1, ty1 = (by1 - 1) / TILE_H + 1;
        const int sx0 = tx0 / BIN_SUPER, sy0 = ty0 / BIN_SUPER;
        const int sw = (tx1 - 1) / BIN_SUPER - sx0 + 1, ns = sw * ((ty1 - 1) / BIN_SUPER - sy0 + 1);
        for (int sb = part * 32; sb < ns; sb += 32 * share) {
            const int si = sb + lane;
            int s_cls = 0, stx = 0, sty = 0;
            if (si < ns) {
                sty = sy0 + si / sw; stx = sx0 + si % sw;
                const int ta = max(tx0, stx * BIN_SUPER), tb = min(tx1, (stx + 1) * BIN_SUPER);
                const int tc = max(ty0, sty * BIN_SUPER), td = min(ty1, (sty + 1) * BIN_SUPER);
                s_cls = quad_tile_class(QE, max(bx0, ta * TILE_W), min(bx1, tb * TILE_W) - 1,
                                        max(by0, tc * TILE_H), min(by1, td * TILE_H) - 1);
            }
            unsigned alive = __ballot_sync(0xffffffffu, s_cls != 0);
            constexpr int PER = BIN_SUPER * BIN_SUPER, GROUPS = 32 / PER;  // surviving blocks expanded per iteration
            while (alive) {   // warp-uniform: the next GROUPS surviving blocks, PER lanes each
                int src = -1;
#pragma unroll
                for (int gidx = 0; gidx < GROUPS; ++gidx) {
                    const int bit = alive ? __ffs(alive) - 1 : -1;
                    alive &= alive - 1;
                    if (lane / PER == gidx) src = bit;
                }
                const int sub = lane % PER;
                const int b_cls = __shfl_sync(0xffffffffu, s_cls, src & 31);
                const int tx = __shfl_sync(0xffffffffu, stx, src & 31) * BIN_SUPER + sub % BIN_SUPER;
                const int ty = __shfl_sync(0xffffffffu, sty, src & 31) * BIN_SUPER + sub / BIN_SUPER;
                int cls = 0, entry = prim, t = 0;
                if (src >= 0 && tx >= tx0 && tx < tx1 && ty >= ty0 && ty < ty1) {
                    const int x0 = max(bx0, tx * TILE_W), x1 = min(bx1, (tx + 1) * TILE_W) - 1;
                    const int y0 = max(by0, ty * TILE_H), y1 = min(by1, (ty + 1) * TILE_H) - 1;
                    cls = b_cls == 2 ? 2 : quad_tile_class(QE, x0, x1, y0, y1);
                    // "full" only counts when the rectangle is the whole tile (clipped to the screen and the band)
                    if (cls == 2 && x0 == tx * TILE_W && x1 == min(Fr.W, (tx + 1) * TILE_W) - 1 &&
                        y0 == max(band_y0, ty * TILE_H) && y1 == min(min(Fr.H, band_y1), (ty + 1) * TILE_H) - 1)
                        entry |= QUAD_FULL_BIT;
                    t = (ty - Fr.tile_row0) * Fr.tiles_x + tx;
                    if (cls) atomicAdd(quad_count + t, 1);
                }
                const unsigned keep = __ballot_sync(0xffffffffu, cls != 0);
                if (cls) my_buf[n_buf + __popc(keep & ((1u << lane) - 1))] = make_int2(entry, t);
                n_buf += __popc(keep);
                if (n_buf > BIN_PAIR_BUF - 32) flush();
            }
        }
    }
    if (n_buf) flush();
}

// exclusive scan of the tile counts of one (view, kind); resets the counts to 0 so k_bin<true> can reuse them
// as cursors.  One CTA of 1024 threads.
__global__ void k_scan(FrameDev Fr, BinDev B, int* __restrict__ host_flags, int* sticky, int view0,
                       const int* __restrict__ sil_count, int sil_cap, int* need_sil) {  // flags: mapped pinned memory
    const int view = blockIdx.x + view0, kind = blockIdx.y;
    const int n_tiles = Fr.tiles_x * Fr.tiles_y;
    int* count = (kind ? B.quad_count : B.tri_count) + (size_t)view * n_tiles;
    int* off = (kind ? B.quad_off : B.tri_off) + (size_t)view * (n_tiles + 1);
    __shared__ int warp_sum[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int base = 0; base < n_tiles; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < n_tiles ? count[i] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) warp_sum[wid] = x;
        __syncthreads();
        if (wid == 0) {
            int s = warp_sum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += y; }
            warp_sum[lane] = s;
        }
        __syncthreads();
        const int prefix = carry + (wid ? warp_sum[wid - 1] : 0) + x - v;
        if (i < n_tiles) { off[i] = prefix; count[i] = 0; }
        __syncthreads();
        if (threadIdx.x == 1023) carry = prefix + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        off[n_tiles] = carry;
        const int cap = kind ? B.quad_cap : B.tri_cap;
        B.overflow[view * 2 + kind] = carry > cap ? carry : 0;
        host_flags[view * 2 + kind] = carry > cap ? carry : 0;  // read by the host after the stream / ticket completes
        // device-resident renders: overflow of ANY call since the last b2r_sync must survive later, fitting calls.
        // Racing views all store a sufficient-or-retried size (the host grows the list and the caller renders again).
        if (sticky && carry > cap) sticky[kind] = carry;
        if (kind == 1 && *sil_count > sil_cap) {   // the silhouette did not fit its records (same for every view)
            *need_sil = *sil_count;
            if (sticky) sticky[3] = *sil_count;
        }
    }
}

// Launch order of the raster CTAs.  A tile under the figure's shadow volume costs ~100x a floor tile; launched in
// screen order, the heavy tiles of the last view start when most of the grid has drained and the launch ends with a
// long, nearly empty tail.  Tiles are therefore bucketed by an estimate of their cost (list lengths) and the raster
// grid walks the buckets heaviest first, the views of the sub-chunk interleaved.  One CTA per view.
constexpr int ORDER_CLASSES = 6;
__device__ __forceinline__ int tile_cost_class(int n_tri, int n_quad, bool full_stencil) {
    // a tile without triangles has no z-buffer: unless the stencil plane itself is wanted, the tile kernel skips it
    const int cost = (n_tri == 0 && !full_stencil) ? 0 : 4 * n_quad + n_tri;
    return cost >= 1024 ? 0 : cost >= 256 ? 1 : cost >= 64 ? 2 : cost >= 16 ? 3 : cost > 0 ? 4 : 5;
}
__global__ void k_order(FrameDev Fr, BinDev B, int view0) {
    const int view = blockIdx.x + view0;
    const int n_tiles = Fr.tiles_x * Fr.tiles_y;
    const int* tri_off = B.tri_off + (size_t)view * (n_tiles + 1);
    const int* quad_off = B.quad_off + (size_t)view * (n_tiles + 1);
    int* order = B.order + (size_t)view * n_tiles;
    __shared__ int cnt[ORDER_CLASSES], base[ORDER_CLASSES];
    if (threadIdx.x < ORDER_CLASSES) cnt[threadIdx.x] = 0;
    __syncthreads();
    for (int t = threadIdx.x; t < n_tiles; t += blockDim.x)
        atomicAdd(&cnt[tile_cost_class(tri_off[t + 1] - tri_off[t], quad_off[t + 1] - quad_off[t], Fr.full_stencil != 0)], 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int c = 0; c < ORDER_CLASSES; ++c) { base[c] = run; run += cnt[c]; cnt[c] = 0; }
        const bool lists_ok = (B.overflow[view * 2] | B.overflow[view * 2 + 1]) == 0;
        B.n_active[view] = lists_ok ? base[ORDER_CLASSES - 1] : 0;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < n_tiles; t += blockDim.x) {
        const int c = tile_cost_class(tri_off[t + 1] - tri_off[t], quad_off[t + 1] - quad_off[t], Fr.full_stencil != 0);
        order[base[c] + atomicAdd(&cnt[c], 1)] = t;
    }
}

// Clip decision of the (screen-filling face, tile) pairs, once per pair, after the lists are filled: one thread per
// (view, noted face, tile of its box).  Where `face_tile_unclipped` proves that no pixel of the tile can fail the clip
// test, the thread finds the pair's entry in the tile's list and clears its TRI_CLIP_BIT: the tile kernel then runs the
// pair without the per-pixel test (three float64 divisions and 24 FMAs per pixel; the floor of the headline scene
// alone is half a million such pixels per view).  A kernel of its own so that the fill pass keeps its 32 registers;
// measured alternatives: inside the tile kernel (one lane per CTA on a ~600-operation dependent chain: slower than no
// elision), inside the fill pass (128 registers, a quarter of the occupancy: fill pass 0.06 -> 0.32 ms).
constexpr int ELIDE_MAX_FACES = 8;   // noted faces per view that get the treatment (the rest keep the per-pixel test)
__global__ void __launch_bounds__(128) k_clip_elide(FrameDev Fr, const TriBox* __restrict__ boxes, const TriRec* __restrict__ tris,
                                                    const double4* __restrict__ pos, const int4* __restrict__ face_vf,
                                                    const ViewDev* __restrict__ views, BinDev B, int view0) {
    const int view = blockIdx.z + view0, h = blockIdx.y;
    if (h >= min(B.huge_count[view], BIN_HUGE_CAP)) return;
    if (B.overflow[view * 2] | B.overflow[view * 2 + 1]) return;
    const int n_tiles = Fr.tiles_x * Fr.tiles_y;
    const int face = B.huge_list[view * BIN_HUGE_CAP + h];
    const TriBox r = boxes[(size_t)view * Fr.n_faces + face];
    if (!(r.flags & TR_NEEDS_CLIP)) return;
    const int bx0 = r.bx0, bx1 = r.bx1, by0 = max((int)r.by0, Fr.row_begin), by1 = min((int)r.by1, Fr.row_end);
    const int tx0 = bx0 / TILE_W, ty0 = by0 / TILE_H, tw = (bx1 - 1) / TILE_W - tx0 + 1;
    const int nt = tw * ((by1 - 1) / TILE_H - ty0 + 1);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nt) return;
    const int tx = tx0 + i % tw, ty = ty0 + i / tw;
    if (!face_tile_unclipped(Fr, tris + (size_t)view * Fr.n_faces, pos, face_vf, views[view], face, tx, ty)) return;
    const int t = (ty - Fr.tile_row0) * Fr.tiles_x + tx;
    const int* tri_off = B.tri_off + (size_t)view * (n_tiles + 1);
    int* tri_list = B.tri_list + (size_t)view * B.tri_cap;
    for (int k = tri_off[t]; k < tri_off[t + 1]; ++k)
        if (tri_list[k] == (face | TRI_CLIP_BIT)) { tri_list[k] = face; break; }
}

// =====================================================================================================================
// tile raster: z -> stencil -> winner, all in shared memory
// =====================================================================================================================
// Relaxed atomic store to shared memory: several lanes / warps may name the same pixel as "last improver" in the same
// round (the verification pass sorts it out), so the store must be an atomic access to be defined behaviour.  Costs the
// same as a plain STS.
__device__ __forceinline__ void store_relaxed_smem(int* p, int v) {
    asm volatile("st.relaxed.cta.shared.b32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}

constexpr int RASTER_WARPS = RASTER_THREADS / 32;
constexpr int STAGE_TRIS = 32;   // triangle records staged in shared memory per round
#ifndef B2R_STAGE_CLIP
#define B2R_STAGE_CLIP 16
#endif
constexpr int STAGE_CLIP = B2R_STAGE_CLIP;   // of which at most this many need the per-pixel clip test (clip coordinates staged too)
constexpr int REC_DOUBLES = 17;  // 128-byte record + 8 bytes of padding: consecutive records start on different banks
static_assert(sizeof(TriRec) == 128, "TriRec layout");

// =====================================================================================================================
// shading
// =====================================================================================================================
// Returns false when the lookup falls outside the map, where NumPy's fancy indexing raises IndexError in the
// reference: an index below -size (the upper side is clipped), or a NaN coordinate (astype(int32) yields INT_MIN).
__device__ __forceinline__ bool texel_fetch(const TextureDev& T, const double P[3], const double uu[3], const double vv[3],
                                            float out[3]) {
    // Face.get_UV (core.py:138-143): clip(max=1), scale by (size-1), truncate, python-style negative wrap
    double cu = gemv3(P[0], P[1], P[2], uu[0], uu[1], uu[2]);
    double cv = gemv3(P[0], P[1], P[2], vv[0], vv[1], vv[2]);
    cu = cu > 1.0 ? 1.0 : cu;
    double rv = 1.0 - cv;
    rv = rv > 1.0 ? 1.0 : rv;
    int col = (int)(cu * (double)(T.width - 1));
    int row = (int)(rv * (double)(T.height - 1));
    if (col < 0) col += T.width;
    if (row < 0) row += T.height;
    bool ok = (cu == cu) && (rv == rv);
    if (col < 0 || col >= T.width) { col = 0; ok = false; }
    if (row < 0 || row >= T.height) { row = 0; ok = false; }
    const uchar4 t = __ldg(T.texels + (size_t)row * T.width + col);
    out[0] = texel_decode(t.x, T.decode); out[1] = texel_decode(t.y, T.decode); out[2] = texel_decode(t.z, T.decode);
    return ok;
}

// normalize() for shading vectors: x * rsqrt(|x|^2).  Within an ulp or two of the reference's x / sqrt(.) -- far
// below what survives the uint8 quantisation (the exact form is kept wherever a comparison depends on it).
__device__ __forceinline__ double fast_rsqrt(double s) {
    // float32 hardware estimate (2^-22) polished by two Newton steps in float64: relative error ~1e-16 like rsqrt(),
    // without its special-case branches; shading vectors are O(1), anything outside float range takes the library path
    if (!(s > 1e-30 && s < 1e30)) return rsqrt(s);
    double y = (double)rsqrtf((float)s);
    const double h = 0.5 * s;
    y = y * fma(-h * y, y, 1.5);
    y = y * fma(-h * y, y, 1.5);
    return y;
}
__device__ __forceinline__ void shade_norm3(double v[3]) {
    const double s = (v[0] * v[0] + v[1] * v[1]) + v[2] * v[2];
    if (s > 0) {
        const double r = fast_rsqrt(s);
        v[0] *= r; v[1] *= r; v[2] *= r;
    }
}

__device__ __forceinline__ double pow_ns(double x, const MaterialDev& M) {
    if (M.ns_log2 >= 0) {  // Ns = 2^k (the default 64, 32, ...): k squarings
        double r = x;
        if (M.ns_log2 == 6) { r *= r; r *= r; r *= r; r *= r; r *= r; r *= r; return r; }   // the reference's default Ns = 64
        for (int i = 0; i < M.ns_log2; ++i) r *= r;
        return r;
    }
    if (M.ns_int >= 0) {  // exponentiation by squaring for any other small integer shininess
        double r = 1.0, b = x;
        int e = M.ns_int;
        while (e) { if (e & 1) r *= b; b *= b; e >>= 1; }
        return r;
    }
    return pow(x, M.Ns);
}

// x ** 0.8 in float32 for the tonemap (core.py:640).  NumPy's float32 power is itself only accurate to about an
// ulp (SVML), so bit parity is not defined here; this version is within ~2 ulp of the true value at a fifth of the
// cost of powf: a hardware log2/exp2 estimate polished by one Newton step on y^5 = x^4.
__device__ __forceinline__ float pow08(float x) {
    if (!(x > 0.0f)) return 0.0f;
    float lg, y;  // bare MUFU.LG2 / MUFU.EX2: frame values are 0 or normal floats in (2^-126, 1]
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(0.8f * lg));
    const float x2 = x * x, t = x2 * x2;
    const float y2 = y * y, y5 = y2 * y2 * y;
    return y * fmaf(0.2f, __fdividef(t, y5), 0.8f);
}

// (c ** 0.8 * 255).astype(uint8) of one pixel, R | G << 8 | B << 16  (core.py:640)
__device__ __forceinline__ unsigned tonemap_pack(const float c[3]) {
    unsigned packed = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float v = __fmul_rn(pow08(c[k]), 255.0f);
        packed |= ((unsigned)(int)v & 0xffu) << (8 * k);
    }
    return packed;
}

// The same with x ** 0.8 evaluated in float64 and rounded once: flat_shading / gouraud write values up to 255 into the
// frame, so the tonemapped value reaches ~21 000 before the uint8 cast wraps it -- two float32 ulp of the fast form above
// would move ~1 % of those pixels across an integer.
__device__ __forceinline__ unsigned tonemap_pack_wide(const float c[3]) {
    unsigned packed = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float p = c[k] > 0.0f ? (float)pow((double)c[k], 0.8) : 0.0f;
        const float v = __fmul_rn(p, 255.0f);
        packed |= ((unsigned)(int)v & 0xffu) << (8 * k);
    }
    return packed;
}

// Small transfers done by the SMs instead of the copy engines.  A cudaMemcpyAsync / cudaMemsetAsync on the compute
// stream queues behind whatever large transfer the same copy engine is busy with (the frames of the previous batch on
// their way to the host): measured, that stalls the whole pipeline by ~0.8 ms per 16-frame batch.
__global__ void k_copy_words(unsigned* __restrict__ dst, const unsigned* __restrict__ src, size_t n) {  // src: mapped pinned host memory
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
__global__ void k_zero_words(unsigned* __restrict__ dst, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = 0u;
}

// Per-call constants evaluated once on the device with the functions the pixel kernels use: the tonemapped
// constant background (core.py:588 `np.full(..., background)` through core.py:640).
__global__ void k_frame_consts(FrameDev Fr, int* __restrict__ counters, int n_counters, unsigned* __restrict__ bg_packed) {
    for (int i = threadIdx.x; i < n_counters; i += blockDim.x) counters[i] = 0;  // silhouette counts of this call
    if (threadIdx.x == 0) *bg_packed = tonemap_pack(Fr.background);
}

__device__ __forceinline__ float clip01(double v) { return (float)(v < 0.05 ? 0.05 : (v > 1.0 ? 1.0 : v)); }

// The flat normal of a face in the vertex dtype: Face.unit_normal_world_space (core.py:127-130)
__device__ __forceinline__ void face_unit_normal(const ShadeStatic& fs, double fn[3]) {
    const double (*wp)[3] = fs.wp;
    if (fs.flags & FS_VTX_F32) {
        float a[3], b[3], c[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) { a[k] = (float)wp[0][k]; b[k] = (float)wp[1][k]; c[k] = (float)wp[2][k]; }
        const float e0x = __fsub_rn(b[0], a[0]), e0y = __fsub_rn(b[1], a[1]), e0z = __fsub_rn(b[2], a[2]);
        const float e1x = __fsub_rn(c[0], a[0]), e1y = __fsub_rn(c[1], a[1]), e1z = __fsub_rn(c[2], a[2]);
        const float cx = __fsub_rn(__fmul_rn(e0y, e1z), __fmul_rn(e0z, e1y));
        const float cy = __fsub_rn(__fmul_rn(e0z, e1x), __fmul_rn(e0x, e1z));
        const float cz = __fsub_rn(__fmul_rn(e0x, e1y), __fmul_rn(e0y, e1x));
        float l = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(cx, cx), __fmul_rn(cy, cy)), __fmul_rn(cz, cz)));
        if (l == 0.0f) l = 1.0f;
        fn[0] = (double)__fdiv_rn(cx, l); fn[1] = (double)__fdiv_rn(cy, l); fn[2] = (double)__fdiv_rn(cz, l);
    } else {
        const double e0x = wp[1][0] - wp[0][0], e0y = wp[1][1] - wp[0][1], e0z = wp[1][2] - wp[0][2];
        const double e1x = wp[2][0] - wp[0][0], e1y = wp[2][1] - wp[0][1], e1z = wp[2][2] - wp[0][2];
        fn[0] = e0y * e1z - e0z * e1y; fn[1] = e0z * e1x - e0x * e1z; fn[2] = e0x * e1y - e0y * e1x;
        normalize3(fn);
    }
}

// flat_shading / gouraud / pbr (triangular.py:174-263), the shading functions the reference keeps commented out at the
// call site of rasterize() (triangular.py:127-130).  They receive the SCREEN barycentrics and -- pbr -- the face's
// screen-space vertices (viewport x, y, linearised z) as positions, exactly what the reference hands them there.
__device__ __noinline__ void shade_alt_pixel(const SceneDev& S, const ViewDev& V, const LightDev& L, const TriRec& r, int face,
                                             int px, int py, int shading, float out[3]) {
    const ShadeStatic& fs = S.shade[face];
    const MaterialDev& M = S.mats[fs.material];
    if (shading == B2R_SHADE_FLAT) {
        double fn[3];
        face_unit_normal(fs, fn);
        double it = seq3(fn[0], fn[1], fn[2], L.direction[0], L.direction[1], L.direction[2]);
        it = it < 0.3 ? 0.3 : (it > 1.0 ? 1.0 : it);
        out[0] = out[1] = out[2] = (float)(it * 255.0);
        return;
    }
    float bu, bv, bw;
    tri_bary(r, px, py, bu, bv, bw);
    const double (*vn)[3] = fs.vn;
    double nb[3];   // bar @ face.normals
    if (fs.flags & FS_NRM_F32) {
#pragma unroll
        for (int k = 0; k < 3; ++k)
            nb[k] = (double)__fmaf_rn(bw, (float)vn[2][k], __fmaf_rn(bv, (float)vn[1][k], __fmul_rn(bu, (float)vn[0][k])));
    } else {
#pragma unroll
        for (int k = 0; k < 3; ++k) nb[k] = seq3((double)bu, (double)bv, (double)bw, vn[0][k], vn[1][k], vn[2][k]);
    }
    if (shading == B2R_SHADE_GOURAUD) {
        double it = (nb[0] * L.direction[0] + nb[1] * L.direction[1]) + nb[2] * L.direction[2];
        it = it < 0.0 ? 0.0 : (it > 1.0 ? 1.0 : it);
        out[0] = out[1] = out[2] = (float)(it * 255.0);
        return;
    }
    // pbr
    const double PI = 3.141592653589793;
    const double metallic = M.Pm, roughness = M.Pr;
    double N[3] = {nb[0], nb[1], nb[2]};
    if (fs.flags & FS_NRM_F32) {   // normalize() on the float32 array
        const float nx = (float)nb[0], ny = (float)nb[1], nz = (float)nb[2];
        float l = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(nx, nx), __fmul_rn(ny, ny)), __fmul_rn(nz, nz)));
        if (l == 0.0f) l = 1.0f;
        N[0] = (double)__fdiv_rn(nx, l); N[1] = (double)__fdiv_rn(ny, l); N[2] = (double)__fdiv_rn(nz, l);
    } else normalize3(N);
    const double b0 = (double)bu, b1 = (double)bv, b2 = (double)bw;
    const double sx[3] = {r.ax, r.ax + r.v0x, r.ax + r.v1x}, sy[3] = {r.ay, r.ay + r.v0y, r.ay + r.v1y};
    const double pos[3] = {seq3(b0, b1, b2, sx[0], sx[1], sx[2]), seq3(b0, b1, b2, sy[0], sy[1], sy[2]),
                           seq3(b0, b1, b2, r.zl[0], r.zl[1], r.zl[2])};
    double Vd[3] = {V.cam_pos[0] - pos[0], V.cam_pos[1] - pos[1], V.cam_pos[2] - pos[2]};
    normalize3(Vd);
    const double F0 = 0.04 * (1.0 - metallic) + 1.0 * metallic;
    double lv[3] = {L.position[0] - pos[0], L.position[1] - pos[1], L.position[2] - pos[2]};
    const double distance = norm3(lv[0], lv[1], lv[2]);
    double Ld[3] = {lv[0], lv[1], lv[2]};
    normalize3(Ld);
    double Hd[3] = {Vd[0] + Ld[0], Vd[1] + Ld[1], Vd[2] + Ld[2]};
    normalize3(Hd);
    const double attenuation = 1.0 / (distance * distance);
    const double a = roughness * roughness, a2 = a * a;
    double NdotH = dot3_plain(N, Hd); NdotH = NdotH < 0 ? 0 : NdotH;
    double den = NdotH * NdotH * (a2 - 1.0) + 1.0;
    den = PI * den * den;
    const double NDF = a2 / den;
    double NdotV = dot3_plain(N, Vd); NdotV = NdotV < 0 ? 0 : NdotV;
    double NdotL = dot3_plain(N, Ld); NdotL = NdotL < 0 ? 0 : NdotL;
    const double r1 = roughness + 1.0, kk = (r1 * r1) / 8.0;
    const double G = (NdotL / (NdotL * (1.0 - kk) + kk)) * (NdotV / (NdotV * (1.0 - kk) + kk));
    double HdotV = dot3_plain(Hd, Vd); HdotV = HdotV < 0 ? 0 : HdotV;
    const double om = 1.0 - HdotV, om2 = om * om;
    const double Fr = F0 + (1.0 - F0) * (om2 * om2 * om);
    const double kD = (1.0 - Fr) * (1.0 - metallic);
    const double specular = (NDF * G * Fr) / (4.0 * NdotV * NdotL + 0.0001);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double radiance = L.color[k] * attenuation;
        const double Lo = (kD * 1.0 / PI + specular) * radiance * NdotL;
        double color = M.Ka[k] + Lo;
        color = color / (color + 1.0);
        out[k] = (float)pow(color, 1.0 / 2.2);
    }
}

// general_shading for one pixel (triangular.py:135-171)
// Returns false when a texture lookup fell outside its map (the reference raises IndexError there).
__device__ bool shade_face_pixel(const SceneDev& S, const ViewDev& V, const LightDev& L, const TriRec& r, int face,
                                 int px, int py, bool lit, float out[3]) {
    bool tex_ok = true;
    const ShadeStatic& fs = S.shade[face];
    const MaterialDev& M = S.mats[fs.material];
    float bu, bv, bw;
    tri_bary(r, px, py, bu, bv, bw);
    double P[3];
    {   // Face.screen_perspective (core.py:155-160); one reciprocal instead of three divisions: the result only
        // feeds shading (<= 1 ulp apart), the clip test in the raster keeps the exact form
        const double b0 = (double)bu, b1 = (double)bv, b2 = (double)bw;
        const double inv_w = 1.0 / gemv3(b0, b1, b2, r.d[0], r.d[1], r.d[2]);
        P[0] = b0 * r.d[0] * inv_w; P[1] = b1 * r.d[1] * inv_w; P[2] = b2 * r.d[2] * inv_w;
    }
    const double* uu = fs.uu;
    const double* vv = fs.vv;
    double albedo[3];
    if (M.map_Kd >= 0) {
        float t[3];
        tex_ok = texel_fetch(S.tex[M.map_Kd], P, uu, vv, t) && tex_ok;
        albedo[0] = (double)t[0]; albedo[1] = (double)t[1]; albedo[2] = (double)t[2];
    } else {
        albedo[0] = M.Kd[0]; albedo[1] = M.Kd[1]; albedo[2] = M.Kd[2];
    }
    const double (*wp)[3] = fs.wp;
    double frag[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) frag[k] = seq3(P[0], P[1], P[2], wp[0][k], wp[1][k], wp[2][k]);
    double dl[3] = {L.position[0] - frag[0], L.position[1] - frag[1], L.position[2] - frag[2]};
    const double dist = norm3(dl[0], dl[1], dl[2]);
    const double att = 1.0 / (L.constant + dist * (L.linear + L.quadratic * dist));  // core.py:517-524
    if (!lit) {
#pragma unroll
        for (int k = 0; k < 3; ++k) out[k] = clip01(att * L.ambient[k] * albedo[k]);
        return tex_ok;
    }
    // Face.get_normals (core.py:175-189)
    const double (*vn)[3] = fs.vn;
    double N[3];
    if (M.norm >= 0) {
        const TextureDev& T = S.tex[M.norm];
        float t[3];
        tex_ok = texel_fetch(T, P, uu, vv, t) && tex_ok;
        if (T.tangent) {  // Face.tangent_ (core.py:191-224)
            double n[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) n[k] = seq3(P[0], P[1], P[2], vn[0][k], vn[1][k], vn[2][k]);
            shade_norm3(n);
            // rows of A (core.py:210-213): r0 = b - a, r1 = c - a (vertex dtype arithmetic), r2 = n
            double r0[3], r1[3];
            if (fs.flags & FS_VTX_F32) {
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    r0[k] = (double)__fsub_rn((float)wp[1][k], (float)wp[0][k]);
                    r1[k] = (double)__fsub_rn((float)wp[2][k], (float)wp[0][k]);
                }
            } else {
#pragma unroll
                for (int k = 0; k < 3; ++k) { r0[k] = wp[1][k] - wp[0][k]; r1[k] = wp[2][k] - wp[0][k]; }
            }
            double du1, du2, dv1, dv2;
            if (fs.flags & FS_UV_F32) {
                du1 = (double)__fsub_rn((float)uu[1], (float)uu[0]); du2 = (double)__fsub_rn((float)uu[2], (float)uu[0]);
                dv1 = (double)__fsub_rn((float)vv[1], (float)vv[0]); dv2 = (double)__fsub_rn((float)vv[2], (float)vv[0]);
            } else {
                du1 = uu[1] - uu[0]; du2 = uu[2] - uu[0]; dv1 = vv[1] - vv[0]; dv2 = vv[2] - vv[0];
            }
            // inv(A) @ (d1, d2, 0) = (d1 * (r1 x n) + d2 * (n x r0)) / det(A); the result is normalised right away
            // (core.py:221-222), so only the SIGN of the determinant survives: no matrix inverse is formed.
            const double c1[3] = {r1[1] * n[2] - r1[2] * n[1], r1[2] * n[0] - r1[0] * n[2], r1[0] * n[1] - r1[1] * n[0]};
            const double c2[3] = {n[1] * r0[2] - n[2] * r0[1], n[2] * r0[0] - n[0] * r0[2], n[0] * r0[1] - n[1] * r0[0]};
            const double det = (r0[0] * c1[0] + r0[1] * c1[1]) + r0[2] * c1[2];
            const double sg = det < 0 ? -1.0 : 1.0;
            double ti[3], tj[3];
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                ti[q] = sg * (du1 * c1[q] + du2 * c2[q]);
                tj[q] = sg * (dv1 * c1[q] + dv2 * c2[q]);
            }
            shade_norm3(ti); shade_norm3(tj);
#pragma unroll
            for (int q = 0; q < 3; ++q) N[q] = seq3(ti[q], tj[q], n[q], (double)t[0], (double)t[1], (double)t[2]);
        } else {
            N[0] = (double)t[0]; N[1] = (double)t[1]; N[2] = (double)t[2];
        }
    } else if (fs.flags & FS_HAS_NORMALS) {
#pragma unroll
        for (int k = 0; k < 3; ++k) N[k] = seq3(P[0], P[1], P[2], vn[0][k], vn[1][k], vn[2][k]);
    } else {  // flat normal in the vertex dtype (core.py:186-187)
        double fn[3];
        face_unit_normal(fs, fn);
#pragma unroll
        for (int k = 0; k < 3; ++k) N[k] = seq3(P[0], P[1], P[2], fn[k], fn[k], fn[k]);
    }
    shade_norm3(N);
    double Ld[3];
    if (L.type == B2R_LIGHT_DIRECTIONAL) { Ld[0] = L.direction[0]; Ld[1] = L.direction[1]; Ld[2] = L.direction[2]; }
    else { Ld[0] = dl[0]; Ld[1] = dl[1]; Ld[2] = dl[2]; shade_norm3(Ld); }
    double Vd[3] = {V.cam_pos[0] - frag[0], V.cam_pos[1] - frag[1], V.cam_pos[2] - frag[2]};
    shade_norm3(Vd);
    if (L.type == B2R_LIGHT_SPOT) {  // triangular.py:157-161, core.py:497-515
        double x = dot3_plain(L.direction, Ld);
        x = (x - L.spot_cos_outer) / (L.spot_cos_inner - L.spot_cos_outer);
        x = x < 0.0 ? 0.0 : (x > 1.0 ? 1.0 : x);
        const double s = x * x * (3.0 - 2.0 * x);
        albedo[0] *= s; albedo[1] *= s; albedo[2] *= s;
    }
    double spec_light[3];
    if (M.map_Ks >= 0) {  // core.py:145-153: float32 texel * 255
        float t[3];
        tex_ok = texel_fetch(S.tex[M.map_Ks], P, uu, vv, t) && tex_ok;
        const double s = (double)__fmul_rn(t[0], 255.0f);
        spec_light[0] = spec_light[1] = spec_light[2] = s;
    } else {
        spec_light[0] = M.Ks255[0]; spec_light[1] = M.Ks255[1]; spec_light[2] = M.Ks255[2];
    }
    double Hd[3] = {Ld[0] + Vd[0], Ld[1] + Vd[1], Ld[2] + Vd[2]};
    shade_norm3(Hd);
    double nh = dot3_plain(N, Hd);
    nh = nh < 0 ? 0 : nh;
    const double sr = pow_ns(nh, M);
    const double nl = dot3_plain(N, Ld);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double specular = L.color[k] * sr * L.specular_strength * spec_light[k];
        const double diffuse = nl * L.color[k];
        out[k] = clip01(att * albedo[k] * (L.ambient[k] + diffuse + specular));
    }
    return tex_ok;
}

// ---- float32 lighting (production) ---------------------------------------------------------------------------------
// general_shading with the float64 arithmetic of the reference kept where a result is discontinuous in its inputs --
// coverage, perspective weights and texel addressing -- and float32 for the lighting sum itself: north_star's bar for
// shaded RGB is 1 LSB on >= 99.9 % of the pixels; float32 lighting moves the tonemapped value by ~1e-4 LSB, i.e. it flips
// the truncated uint8 of about one channel in 10^4 (measured per fixture in profiles/), never by more than one.  The
// float64 form above stays available (B2R_SHADE_F64=1) and is what the fused / debug paths use.
__device__ __forceinline__ float rsqrt_fast(float s) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(s)); return y; }
__device__ __forceinline__ float rcp_fast(float s) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(s)); return y; }
__device__ __forceinline__ float dot3f(const float a[3], const float b[3]) { return fmaf(a[2], b[2], fmaf(a[1], b[1], a[0] * b[0])); }
__device__ __forceinline__ void norm3f(float v[3]) {
    const float s = dot3f(v, v);
    if (s > 0.f) { const float r = rsqrt_fast(s); v[0] *= r; v[1] *= r; v[2] *= r; }
}
__device__ __forceinline__ float clip01f(float v) { return fminf(fmaxf(v, 0.05f), 1.0f); }
// texel address of Face.get_UV (core.py:138-143) in float64, the texel itself undecoded
__device__ __forceinline__ bool texel_raw(const TextureDev& T, const double P[3], const double uu[3], const double vv[3], uchar4& t) {
    double cu = gemv3(P[0], P[1], P[2], uu[0], uu[1], uu[2]);
    double cv = gemv3(P[0], P[1], P[2], vv[0], vv[1], vv[2]);
    cu = cu > 1.0 ? 1.0 : cu;
    double rv = 1.0 - cv;
    rv = rv > 1.0 ? 1.0 : rv;
    int col = (int)(cu * (double)(T.width - 1));
    int row = (int)(rv * (double)(T.height - 1));
    if (col < 0) col += T.width;
    if (row < 0) row += T.height;
    bool ok = (cu == cu) && (rv == rv);
    if (col < 0 || col >= T.width) { col = 0; ok = false; }
    if (row < 0 || row >= T.height) { row = 0; ok = false; }
    t = __ldg(T.texels + (size_t)row * T.width + col);
    return ok;
}
__device__ __forceinline__ void texel_decode_f(const uchar4 t, int snorm, float out[3]) {
    const float a = snorm ? (2.0f / 255.0f) : (1.0f / 255.0f), b = snorm ? -1.0f : 0.0f;
    out[0] = fmaf((float)t.x, a, b); out[1] = fmaf((float)t.y, a, b); out[2] = fmaf((float)t.z, a, b);
}

__device__ bool shade_face_pixel_f32(const SceneDev& S, const ViewDev& V, const LightLite& L, int light_type, const TriRec& r,
                                     int face, int px, int py, bool lit, float out[3]) {
    bool tex_ok = true;
    const ShadeLite& fs = S.shade_lite[face];
    const MaterialDev& M = S.mats[fs.material];
    float bu, bv, bw;
    tri_bary(r, px, py, bu, bv, bw);
    double P[3];
    {   // Face.screen_perspective (core.py:155-160) in float64: it feeds the texel addresses
        const double b0 = (double)bu, b1 = (double)bv, b2 = (double)bw;
        const double inv_w = 1.0 / gemv3(b0, b1, b2, r.d[0], r.d[1], r.d[2]);
        P[0] = b0 * r.d[0] * inv_w; P[1] = b1 * r.d[1] * inv_w; P[2] = b2 * r.d[2] * inv_w;
    }
    const float p0 = (float)P[0], p1 = (float)P[1], p2 = (float)P[2];
    float albedo[3];
    if (M.map_Kd >= 0) {
        const TextureDev& T = S.tex[M.map_Kd];
        uchar4 t;
        tex_ok = texel_raw(T, P, fs.uu, fs.vv, t) && tex_ok;
        texel_decode_f(t, T.decode, albedo);
    } else {
        albedo[0] = M.Kdf[0]; albedo[1] = M.Kdf[1]; albedo[2] = M.Kdf[2];
    }
    float frag[3], dl[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        frag[k] = fmaf(p2, fs.wp[2][k], fmaf(p1, fs.wp[1][k], p0 * fs.wp[0][k]));
        dl[k] = L.position[k] - frag[k];
    }
    const float d2 = dot3f(dl, dl);
    const float rd = d2 > 0.f ? rsqrt_fast(d2) : 0.f;
    const float dist = d2 * rd;
    const float att = rcp_fast(fmaf(dist, fmaf(L.quadratic, dist, L.linear), L.constant));  // core.py:517-524
    if (!lit) {
#pragma unroll
        for (int k = 0; k < 3; ++k) out[k] = clip01f(att * L.ambient[k] * albedo[k]);
        return tex_ok;
    }
    // Face.get_normals (core.py:175-189)
    float N[3];
    if (M.norm >= 0) {
        const TextureDev& T = S.tex[M.norm];
        uchar4 tq;
        tex_ok = texel_raw(T, P, fs.uu, fs.vv, tq) && tex_ok;
        float t[3];
        texel_decode_f(tq, T.decode, t);
        if (T.tangent) {  // Face.tangent_ (core.py:191-224)
            float n[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) n[k] = fmaf(p2, fs.vn[2][k], fmaf(p1, fs.vn[1][k], p0 * fs.vn[0][k]));
            norm3f(n);
            const float* r0 = fs.r0;
            const float* r1 = fs.r1;
            // inv(A) @ (d1, d2, 0) = (d1 * (r1 x n) + d2 * (n x r0)) / det(A), normalised right away: only the sign of det survives
            const float c1[3] = {fmaf(r1[1], n[2], -(r1[2] * n[1])), fmaf(r1[2], n[0], -(r1[0] * n[2])), fmaf(r1[0], n[1], -(r1[1] * n[0]))};
            const float c2[3] = {fmaf(n[1], r0[2], -(n[2] * r0[1])), fmaf(n[2], r0[0], -(n[0] * r0[2])), fmaf(n[0], r0[1], -(n[1] * r0[0]))};
            const float det = dot3f(r0, c1);
            const float sg = det < 0.f ? -1.0f : 1.0f;
            float ti[3], tj[3];
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                ti[q] = sg * fmaf(fs.du1, c1[q], fs.du2 * c2[q]);
                tj[q] = sg * fmaf(fs.dv1, c1[q], fs.dv2 * c2[q]);
            }
            norm3f(ti); norm3f(tj);
#pragma unroll
            for (int q = 0; q < 3; ++q) N[q] = fmaf(n[q], t[2], fmaf(tj[q], t[1], ti[q] * t[0]));
        } else {
            N[0] = t[0]; N[1] = t[1]; N[2] = t[2];
        }
    } else {  // interpolated vertex normals, or the flat normal (stored three times: core.py:186-187)
#pragma unroll
        for (int k = 0; k < 3; ++k) N[k] = fmaf(p2, fs.vn[2][k], fmaf(p1, fs.vn[1][k], p0 * fs.vn[0][k]));
    }
    norm3f(N);
    float Ld[3];
    if (light_type == B2R_LIGHT_DIRECTIONAL) { Ld[0] = L.direction[0]; Ld[1] = L.direction[1]; Ld[2] = L.direction[2]; }
    else { Ld[0] = dl[0] * rd; Ld[1] = dl[1] * rd; Ld[2] = dl[2] * rd; }
    float Vd[3] = {V.cam_posf[0] - frag[0], V.cam_posf[1] - frag[1], V.cam_posf[2] - frag[2]};
    norm3f(Vd);
    if (light_type == B2R_LIGHT_SPOT) {  // triangular.py:157-161, core.py:497-515
        float x = (dot3f(L.direction, Ld) - L.spot_cos_outer) * L.spot_inv_range;
        x = fminf(fmaxf(x, 0.f), 1.f);
        const float sm = x * x * fmaf(-2.0f, x, 3.0f);
        albedo[0] *= sm; albedo[1] *= sm; albedo[2] *= sm;
    }
    float spec_light[3];
    if (M.map_Ks >= 0) {  // core.py:145-153: float32 texel * 255
        const TextureDev& T = S.tex[M.map_Ks];
        uchar4 tq;
        tex_ok = texel_raw(T, P, fs.uu, fs.vv, tq) && tex_ok;
        float t[3];
        texel_decode_f(tq, T.decode, t);
        spec_light[0] = spec_light[1] = spec_light[2] = t[0] * 255.0f;
    } else {
        spec_light[0] = M.Ks255f[0]; spec_light[1] = M.Ks255f[1]; spec_light[2] = M.Ks255f[2];
    }
    float Hd[3] = {Ld[0] + Vd[0], Ld[1] + Vd[1], Ld[2] + Vd[2]};
    norm3f(Hd);
    const float nh = fmaxf(dot3f(N, Hd), 0.f);
    float sr;
    if (M.ns_log2 == 6) { sr = nh * nh; sr *= sr; sr *= sr; sr *= sr; sr *= sr; sr *= sr; }
    else if (M.ns_log2 >= 0) { sr = nh; for (int i = 0; i < M.ns_log2; ++i) sr *= sr; }
    else sr = (float)pow_ns((double)nh, M);
    const float nl = dot3f(N, Ld);
    const float ss = sr * L.specular_strength;
#pragma unroll
    for (int k = 0; k < 3; ++k)
        out[k] = clip01f(att * albedo[k] * fmaf(L.color[k], fmaf(ss, spec_light[k], nl), L.ambient[k]));
    return tex_ok;
}

// cube_map.py:63-101 for one background pixel; returns false when neither screen triangle covers it
__device__ __forceinline__ bool skybox_pixel(const SceneDev& S, const ViewDev& V, int sky_size, int px, int py, float out[3]) {
#pragma unroll
    for (int t = 1; t >= 0; --t) {  // the second triangle is written last in the reference, so it wins
        const SkyTri& T = V.sky[t];
        if (!T.ok) continue;
        const long long v2x = px - T.ax, v2y = py - T.ay;
        const float d20 = (float)(v2x * T.v0x + v2y * T.v0y), d21 = (float)(v2x * T.v1x + v2y * T.v1y);
        const float bv = __fmul_rn(__fsub_rn(__fmul_rn(T.d11, d20), __fmul_rn(T.d01, d21)), T.inv);
        const float bw = __fmul_rn(__fsub_rn(__fmul_rn(T.d00, d21), __fmul_rn(T.d01, d20)), T.inv);
        const float bu = __fsub_rn(__fsub_rn(1.0f, bv), bw);
        if (!(bu >= 0.0f && bv >= 0.0f && bw >= 0.0f)) continue;
        double r[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) r[k] = seq3((double)bu, (double)bv, (double)bw, T.rays[0][k], T.rays[1][k], T.rays[2][k]);
        int axis = 0;
        double best = fabs(r[0]);
        if (fabs(r[1]) > best) { best = fabs(r[1]); axis = 1; }
        if (fabs(r[2]) > best) { axis = 2; }
        const double amp = axis == 0 ? r[0] : (axis == 1 ? r[1] : r[2]);
        const double o0 = axis == 0 ? r[1] : r[0], o1 = axis == 2 ? r[1] : r[2];
        const double u0 = (o0 / amp + 1.0) / 2.0, u1 = (o1 / amp + 1.0) / 2.0;
        const int side = (amp < 0 ? 1 : 0) + axis * 2;
        long long i0 = (long long)(u0 * (double)sky_size - 1.0), i1 = (long long)(u1 * (double)sky_size - 1.0);
        if (i0 < 0) i0 += sky_size;
        if (i1 < 0) i1 += sky_size;
        if (i0 < 0 || i0 >= sky_size) i0 = 0;
        if (i1 < 0 || i1 >= sky_size) i1 = 0;
        const uchar4 tx = __ldg(S.sky + ((size_t)side * sky_size + i0) * sky_size + i1);
        out[0] = texel_decode(tx.x, 0); out[1] = texel_decode(tx.y, 0); out[2] = texel_decode(tx.z, 0);
        return true;
    }
    return false;
}

#ifndef B2R_SHADE_MINB
#define B2R_SHADE_MINB 12
#endif

// =====================================================================================================================
// k_tile<FUSED>: depth -> stencil -> winner (-> SHADING when FUSED) of one 32x32 tile; z-buffer, winner and stencil count
// live in shared memory (swizzled: pixel (x, y) sits at (y << 5) | (x ^ y), so lanes sharing a column do not pile up on
// one bank).
//   FUSED = false (production): the kernel ends by writing ONE packed word per pixel (winner | lit << 31) for
//     k_shade_packed; tiles without primitives write nothing.  Round 1 wrote winner (4 B) + stencil (2 B) for every
//     pixel of the screen.
//   FUSED = true (B2R_FUSED=1, kept as the measured alternative): shading runs at the end of this kernel straight from
//     the shared planes, nothing but the finished pixels reaches HBM.  Measured on the B200 (diablo, 64 views): 5.19 ms
//     against 2.66 + 1.55 ms for the two-kernel form -- the fused kernel's instruction footprint (~100 KB against a
//     32 KB L1.5 instruction cache, CTAs of one SM in different phases) and its 64 registers / 8 CTAs per SM cost more
//     than the 4 bytes per covered pixel of HBM traffic it saves.
// Exactness of the shortcuts: the quad edge function and the quad plane depth, as the reference
// rounds them, are compositions of monotone roundings in px and py, so their extremes over a pixel rectangle sit
// exactly at its corners -- depth-range rejection / acceptance of a (quad, tile) pair, the per-row span search and the
// edge skipping below are exact, not merely conservative.
// Also tried this round and measured slower, hence not kept: (triangle, row, 8-pixel segment) items with incremental
// barycentrics in the depth pass and lane-owned rows / columns or 4-pixel segments in the stencil pass -- fewer
// instructions per pixel, but 25 instead of 30 active lanes per instruction and a longer closing barrier (2.94 ms
// against 2.66 ms for the dense (primitive, pixel) dealing below); rounds of 62 staged triangles in the depth pass (the
// second 30 records in the memory of the still unused stencil plane): +3 % on diablo, no gain on the 1M-triangle torus,
// whose tile time is its 17.8 k shadow quads, not the rounds.
// =====================================================================================================================
// 64-bit warp-wide minimum / maximum through two 32-bit REDUX steps (high words, then the low words of the lanes that
// hold the winning high word) instead of five shuffle rounds of 64-bit values
__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v) {
    const unsigned hi = (unsigned)(v >> 32), lo = (unsigned)v;
    const unsigned mh = __reduce_min_sync(0xffffffffu, hi);
    const unsigned ml = __reduce_min_sync(0xffffffffu, hi == mh ? lo : 0xffffffffu);
    return ((unsigned long long)mh << 32) | ml;
}
__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
    const unsigned hi = (unsigned)(v >> 32), lo = (unsigned)v;
    const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
    const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
    return ((unsigned long long)mh << 32) | ml;
}
__device__ __forceinline__ int tpix(int x, int y) { return (y << 5) | ((x ^ y) & 31); }  // x, y in [0, 32)

constexpr int STAMP_RACED = 0x7fffffff;     // depth pass: two improvements of this pixel in one round (see tile_tris)
constexpr unsigned PACKED_LIT = 0x80000000u;   // bit 31 of the packed winner word: stencil == 0
constexpr unsigned PACKED_NONE = 0x7fffffffu;  // no face (background)

struct TileOut {
    uint8_t* rgb;          // FUSED: (views, H, W, 3) final image rows
    float* f32;            // FUSED, optional: float frame before tonemap, buffer rows
    const unsigned* bg_packed;
    unsigned* packed;      // !FUSED: (views, H, W) winner | lit << 31, buffer rows (background tiles are NOT written:
                           // k_shade_packed recognises them from the empty tile lists)
    int* winner;           // optional debug planes (views, H, W), buffer rows
    short* stencil;
    double* z;
    uint8_t* status;       // optional (views, F)
    // SPLIT launches (a handful of views: every tile's pair list is cut over n_parts CTAs, see k_tile)
    int* split_st;         // (views of the batch, n_tiles, n_parts, 1024) partial stencil counts
    int* split_ticket;     // (views of the batch, n_tiles) parts that have delivered, zero before the launch
    int n_parts;
};

struct TileSmem {
    unsigned long long z[TILE_PX];          // order-preserving keys of the float64 z-buffer (swizzled)   8 KB
    int id[TILE_PX];                        // winner face                                                4 KB
    int st[TILE_PX];                        // stencil count                                              4 KB
    double tri[STAGE_TRIS][REC_DOUBLES];    // staged per-tile triangle list                              4.25 KB
    double clip[STAGE_CLIP][CLIP_DOUBLES];  // clip coordinates of the staged triangles that need them    3 KB
    int face[STAGE_TRIS];
    int start[STAGE_TRIS + 1];              // exclusive scan of the pixel counts of the staged triangles
    int geo[STAGE_TRIS];                    // box inside the tile: x0 | y0 << 8 | width << 16
    signed char clip_slot[STAGE_TRIS];
    unsigned long long red_min[RASTER_WARPS], red_max[RASTER_WARPS];
    int n_round;
    int next_quad;
    int uniform;
    int need_full;
    int use_diff;
    int ticket;
};
// One pass over the tile's triangle list.  PASS 1: zbuf + last improver, PASS 3: full winner pass (+ status bits).
template <int PASS>
__device__ __noinline__ void tile_tris(TileSmem& sm, const double4* __restrict__ pos, const int4* __restrict__ face_vf,
                                       const ViewDev& V, const TriRec* __restrict__ vtris, const int* __restrict__ tri_list,
                                       int t_beg, int t_end, int X0, int Y0, int X1, int Yb0, int Y1, bool rh,
                                       uint8_t* status_view) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int n = 0, round_stamp = 0;
    for (int base = t_beg; base < t_end; base += n) {
        ++round_stamp;
        __syncthreads();  // previous round fully consumed
        if (wid == 0) {   // choose the round: up to 32 triangles, at most STAGE_CLIP of them with a clip test (the tile-list
                          // entry carries that flag: no dependent load of the record here)
            const int cand = min(STAGE_TRIS, t_end - base);
            const int e0 = lane < cand ? tri_list[base + lane] : 0;
            const bool clip = (e0 & TRI_CLIP_BIT) != 0;
            const unsigned mask = __ballot_sync(0xffffffffu, clip);
            const int take = __popc(mask) > STAGE_CLIP ? (int)__fns(mask, 0, STAGE_CLIP + 1) : cand;
            if (lane < take) {
                sm.face[lane] = e0 & (TRI_CLIP_BIT - 1);
                sm.clip_slot[lane] = clip ? (signed char)__popc(mask & ((1u << lane) - 1)) : (signed char)-1;
            }
            if (lane == 0) sm.n_round = take;
        }
        __syncthreads();
        n = sm.n_round;
        for (int u = threadIdx.x; u < n * 16; u += RASTER_THREADS) {  // records: coalesced 8-byte pieces
            const int t = u >> 4, part = u & 15;
            reinterpret_cast<unsigned long long*>(sm.tri[t])[part] =
                __ldg(reinterpret_cast<const unsigned long long*>(vtris + sm.face[t]) + part);
        }
        for (int u = threadIdx.x; u < n * CLIP_DOUBLES; u += RASTER_THREADS) {  // clip coordinates, 4 FMAs each
            const int t = u / CLIP_DOUBLES, j = u - t * CLIP_DOUBLES;
            const int slot = sm.clip_slot[t];
            if (slot < 0) continue;
            const int cam = j / 12, vtx = (j % 12) >> 2, k = j & 3;
            const int4 fv = face_vf[sm.face[t]];
            const double4 p = pos[vtx == 0 ? fv.x : (vtx == 1 ? fv.y : fv.z)];
            const double* M = cam ? V.mvp_dbg : V.mvp;
            sm.clip[slot][j] = fma(p.w, M[12 + k], fma(p.z, M[8 + k], fma(p.y, M[4 + k], p.x * M[k])));
        }
        __syncthreads();
        if (wid == 0) {   // boxes inside the tile and the exclusive scan of their pixel counts
            int npx = 0;
            if (lane < n) {
                const TriRec& r = *reinterpret_cast<const TriRec*>(sm.tri[lane]);
                const int x0 = max((int)r.bx0, X0), x1 = min((int)r.bx1, X1), y0 = max((int)r.by0, Yb0), y1 = min((int)r.by1, Y1);
                const int w = max(x1 - x0, 0), h = max(y1 - y0, 0);
                npx = w * h;
                if (npx >= BIG_BOX_PX && tri_misses_rect(r, x0, x1 - 1, y0, y1 - 1)) npx = 0;  // provably no covered pixel
                sm.geo[lane] = (x0 - X0) | ((y0 - Y0) << 8) | (w << 16);
            }
            int incl = npx;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
            sm.start[lane + 1] = incl;
            if (lane == 0) sm.start[0] = 0;
        }
        __syncthreads();
        // the (triangle, pixel) pairs of the round, flattened into one dense list and dealt 128 at a time: a warp is full
        // whether the tile holds a thousand one-pixel triangles or a single triangle covering all of it
        const int total = sm.start[n];
        for (int k0 = wid * 32; k0 < total; k0 += RASTER_THREADS) {
            const int k = k0 + lane;
            if (k >= total) continue;
            int t = 0;  // largest t with start[t] <= k
#pragma unroll
            for (int step = 16; step; step >>= 1) if (t + step < n && sm.start[t + step] <= k) t += step;
            const int i = k - sm.start[t], geo = sm.geo[t];
            const int w = geo >> 16;
            // i / w for i < 1024, w <= 32: (i + 0.5) / w stays >= 1/64 away from an integer
            const int yy = __float2int_rd(((float)i + 0.5f) * (1.0f / (float)w));
            const int lx = (geo & 0xff) + i - yy * w, ly = ((geo >> 8) & 0xff) + yy;
            const int px = X0 + lx, py = Y0 + ly;
            const TriRec& r = *reinterpret_cast<const TriRec*>(sm.tri[t]);
            const int slot = sm.clip_slot[t];
            const double* cc = sm.clip[slot < 0 ? 0 : slot];
            float bu, bv, bw;
            if (PASS == 1) B2R_STAT(12, 1);
            if (PASS == 1 && slot >= 0) B2R_STAT(10, 1);
            if (!tri_pixel_in(r, cc, slot >= 0, px, py, bu, bv, bw)) continue;
            if (PASS == 1) B2R_STAT(13, 1);
            const double b0 = (double)bu, b1 = (double)bv, b2 = (double)bw;
            const double z = (r.flags & TR_COV_ONE) ? seq3(b0, b1, b2, r.zl[0], r.zl[1], r.zl[2])
                                                    : gemv3(b0, b1, b2, r.zl[0], r.zl[1], r.zl[2]);
            unsigned bits = 1;
            const int face = sm.face[t];
            if (z == z) {
                const int p = tpix(lx, ly);
                const unsigned long long key = zkey(z);
                if (PASS == 1) {
                    // a face of a Model(depth_test=False) never writes z (triangular.py:117); it may still colour the
                    // pixel, which only the full winner pass resolves
                    if (r.flags & TR_NO_ZWRITE) { sm.need_full = 1; continue; }
                    const unsigned long long old = rh ? atomicMin(&sm.z[p], key) : atomicMax(&sm.z[p], key);
                    if (old == key) sm.need_full = 1;                                               // exact tie
                    else if (rh ? (key < old) : (key > old)) {
                        store_relaxed_smem(&sm.id[p], face);   // last improver
                        // Only two improvements of one pixel within the SAME round can leave a stale id behind (their
                        // stores are unordered; rounds are separated by barriers).  The idle stencil plane keeps the
                        // round of a pixel's last improvement; a second one in that round marks the pixel for the
                        // verification below -- every other pixel is trusted as it is.
                        const int prev = atomicMax(&sm.st[p], round_stamp);
                        if (prev == round_stamp) atomicMax(&sm.st[p], STAMP_RACED);
                    }
                } else {
                    // writing faces colour where they ARE the z-buffer; non-writing ones wherever they pass the test
                    // against the final z-buffer (zbuf >= z for RH, <= for LH)
                    const unsigned long long kb = sm.z[p];
                    const bool pass = (r.flags & TR_NO_ZWRITE) ? (rh ? (kb >= key) : (kb <= key)) : (key == kb);
                    if (pass) {
                        atomicMax(&sm.id[p], face);
                        bits |= 2 | (sm.st[p] == 0 ? 4 : 0);
                    }
                }
            }
            if (PASS == 3 && status_view) {
                uint8_t* sp = status_view + face;
                unsigned* wp = (unsigned*)((uintptr_t)sp & ~(uintptr_t)3);
                atomicOr(wp, bits << (8 * ((uintptr_t)sp & 3)));
            }
        }
    }
}

// One 32-pixel row segment of finished pixels -> 96 bytes of the uint8 frame (row H-1-py), packed into 24 word stores.
__device__ __forceinline__ void store_row(const FrameDev& Fr, uint8_t* __restrict__ out_rgb, int view, int x_base, int py,
                                          unsigned packed, bool valid, int lane) {
    uint8_t* row = out_rgb + ((size_t)view * Fr.H + (size_t)(Fr.H - 1 - py)) * Fr.W * 3;
    if (x_base + 32 <= Fr.W && (((size_t)Fr.W * 3) & 3) == 0) {
        const int k = (4 * lane) / 3, s = 8 * (lane - 3 * (lane / 3));
        const unsigned a = __shfl_sync(0xffffffffu, packed, k & 31), b = __shfl_sync(0xffffffffu, packed, (k + 1) & 31);
        const unsigned word = (a >> s) | (b << (24 - s));
        if (lane < 24) reinterpret_cast<unsigned*>(row + (size_t)x_base * 3)[lane] = word;
    } else if (valid) {
        const int px = x_base + lane;
        row[(size_t)px * 3 + 0] = (uint8_t)(packed & 0xff);
        row[(size_t)px * 3 + 1] = (uint8_t)((packed >> 8) & 0xff);
        row[(size_t)px * 3 + 2] = (uint8_t)((packed >> 16) & 0xff);
    }
}

#ifndef B2R_TILE_MINB
#define B2R_TILE_MINB 8
#endif
#ifndef B2R_TILE_MINB_UNFUSED
#define B2R_TILE_MINB_UNFUSED 9
#endif
// SPLIT = true (launches of one or two views, i.e. plain scene.render() calls): such a launch is latency bound -- all of
// its ~1000 active tiles are resident at once and it lasts as long as the heaviest tile under the figure's shadow volume,
// one CTA walking ~150 (quad, tile) pairs (0.30 ms against 0.028 ms per view in a 64-view launch).  Every tile then gets
// n_parts CTAs (blockIdx.z): each repeats the (cheap) depth pass and takes one contiguous part of the tile's pair list;
// the parts leave their partial stencil counts in global memory, and the LAST one to arrive (a ticket per tile) sums them
// and finishes the tile.  A separate instantiation: the batch kernel carries none of it.
template <bool FUSED, bool SPLIT = false>
__global__ void __launch_bounds__(RASTER_THREADS, FUSED ? B2R_TILE_MINB : B2R_TILE_MINB_UNFUSED)
k_tile(SceneDev S, const ViewDev* __restrict__ views, FrameDev Fr, const TriRec* __restrict__ tris,
       const QuadRec* __restrict__ quads, int quad_stride, BinDev B, TileOut O, int view0, int n_sub) {
    __shared__ TileSmem sm;
    const int part = SPLIT ? (int)blockIdx.z : 0, n_parts = SPLIT ? O.n_parts : 1;
    // grid (views of the sub-chunk, tile rank): x runs fastest, so the views stay interleaved within a cost class
    const int view = (int)blockIdx.x + view0;
    // ranks past the active tiles of this view (about half the screen in the headline scene): nothing to do, and
    // k_shade_packed recognises those tiles itself -- leave before any other work
    if (!FUSED && !(O.winner || O.stencil || O.z) && (int)blockIdx.y >= B.n_active[view]) return;
    const ViewDev& V = views[view];
    const int n_tiles = Fr.tiles_x * Fr.tiles_y;
    const int tile = B.order[(size_t)view * n_tiles + blockIdx.y];
    const int tx = tile % Fr.tiles_x, ty = tile / Fr.tiles_x + Fr.tile_row0;
    const int X0 = tx * TILE_W, Y0 = ty * TILE_H;
    const int X1 = min(X0 + TILE_W, Fr.W), Y1 = min(min(Y0 + TILE_H, Fr.H), Fr.row_end);
    const int Yb0 = max(Y0, Fr.row_begin);
    const bool rh = V.system == 1;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int* tri_off = B.tri_off + (size_t)view * (n_tiles + 1);
    const int* quad_off = B.quad_off + (size_t)view * (n_tiles + 1);
    const bool lists_ok = (B.overflow[view * 2] | B.overflow[view * 2 + 1]) == 0;
    const int t_beg = lists_ok ? tri_off[tile] : 0, t_end = lists_ok ? tri_off[tile + 1] : 0;
    const int q_beg = lists_ok ? quad_off[tile] : 0, q_end = lists_ok ? quad_off[tile + 1] : 0;
    const size_t plane = (size_t)view * Fr.H * Fr.W;
    const double z_bg = rh ? __longlong_as_double(0x7FF0000000000000ll) : __longlong_as_double(0xFFF0000000000000ll);

    const bool solo = SPLIT && q_beg == q_end;   // no shadow quads over this tile: nothing to split, part 0 does it alone
    if (solo && part != 0) return;
    if (t_beg == t_end && (q_beg == q_end || !Fr.full_stencil)) {
        // no face can win here: background tile
        if (SPLIT && part != 0) return;
        if (!FUSED && !(O.winner || O.stencil || O.z)) return;   // k_shade_packed sees the empty lists itself
        const unsigned bg = (!FUSED || Fr.bg_mode == B2R_BG_CUBEMAP) ? 0u : __ldg(O.bg_packed);
        for (int row = wid; row < TILE_H; row += RASTER_WARPS) {
            const int py = Y0 + row, px = X0 + lane;
            if (py < Yb0 || py >= Y1) continue;
            unsigned packed = bg;
            float c[3] = {Fr.background[0], Fr.background[1], Fr.background[2]};
            if (FUSED && Fr.bg_mode == B2R_BG_CUBEMAP) {
                c[0] = c[1] = c[2] = 0.f;
                if (px < X1) skybox_pixel(S, V, Fr.sky_size, px, py, c);
                packed = tonemap_pack(c);
            }
            if (px < X1) {
                const size_t g = plane + (size_t)py * Fr.W + px;
                if (FUSED && O.f32) { float* o = O.f32 + g * 3; o[0] = c[0]; o[1] = c[1]; o[2] = c[2]; }
                if (O.winner) O.winner[g] = -1;
                if (O.stencil) O.stencil[g] = 0;
                if (O.z) O.z[g] = z_bg;
            }
            if (FUSED) store_row(Fr, O.rgb, view, X0, py, packed, px < X1, lane);
        }
        return;
    }

    if (threadIdx.x == 0) { B2R_STAT(11, 1); B2R_STAT(14, q_end - q_beg); B2R_STAT(15, t_end - t_beg); }
    const unsigned long long z_init = zkey(z_bg);
    for (int i = threadIdx.x; i < TILE_PX; i += RASTER_THREADS) { sm.z[i] = z_init; sm.id[i] = -1; sm.st[i] = 0; }
    if (threadIdx.x == 0) { sm.uniform = 0; sm.need_full = 0; sm.next_quad = 0; sm.use_diff = 0; }
    const TriRec* vtris = tris + (size_t)view * Fr.n_faces;
    const int* tri_list = B.tri_list + (size_t)view * B.tri_cap;
    uint8_t* status_view = O.status ? O.status + (size_t)view * Fr.n_faces : nullptr;

    if (!(B2R_SKIP & 1)) tile_tris<1>(sm, S.pos, S.face_vf, V, vtris, tri_list, t_beg, t_end, X0, Y0, X1, Yb0, Y1, rh, status_view);
    __syncthreads();
    // ---- winner: the last improver, verified where two improvements raced; full pass on ties / lost races ----
    for (int p = threadIdx.x; p < TILE_PX; p += RASTER_THREADS) {
        const int stamp = sm.st[p];
        sm.st[p] = 0;   // the plane becomes the stencil count
        const int f = sm.id[p];
        const unsigned long long kb = sm.z[p];
        if (f < 0) { if (kb != z_init) sm.need_full = 1; continue; }
        if (stamp != STAMP_RACED || (B2R_SKIP & 4)) continue;
        const int ly = p >> 5, lx = (p ^ ly) & 31;
        const TriRec& r = vtris[f];
        float bu, bv, bw;
        tri_bary(r, X0 + lx, Y0 + ly, bu, bv, bw);
        const double b0 = (double)bu, b1 = (double)bv, b2 = (double)bw;
        const double z = (r.flags & TR_COV_ONE) ? seq3(b0, b1, b2, r.zl[0], r.zl[1], r.zl[2])
                                                : gemv3(b0, b1, b2, r.zl[0], r.zl[1], r.zl[2]);
        if (!(z == z) || zkey(z) != kb) sm.need_full = 1;
    }
    __syncthreads();

    // ---- stencil (triangular.py:341-368) ----
    const bool skip_bg = !Fr.full_stencil;
    // the staging area of the triangle rounds is idle during the stencil phase: 1 KB of it per warp holds the row table
    static_assert(sizeof(sm.tri) >= RASTER_WARPS * TILE_PX, "row tables alias the triangle staging area");
    unsigned char* const rowtab = reinterpret_cast<unsigned char*>(sm.tri) + wid * TILE_PX;
    unsigned long long kb_min = ~0ull, kb_max = 0ull;
#if B2R_ROWDIFF
    // Also idle during the stencil phase: the clip-coordinate staging area.  It holds (a) the z-buffer range of the
    // covered pixels of every tile ROW and (b) a per-row DIFFERENCE ARRAY of stencil increments: a row of a (quad, tile)
    // pair whose span passes the depth test as a whole adds +-1 at the span's first pixel and -+1 behind its last one --
    // two shared-memory atomics instead of one per pixel -- and one prefix sum per row at the end of the phase turns
    // the array into counts.  Entries are 16-bit halves of a word, biased by 0x8000 so that a decrement of the low half
    // never borrows from the high half (a tile sees far fewer than 32 767 pairs); the row pitch of 17 words keeps the
    // rows on different banks.
    constexpr int DIFF_PITCH = 17;
    static_assert(sizeof(sm.clip) >= TILE_H * DIFF_PITCH * 4 + TILE_H * 16, "row arrays alias the clip staging area");
    unsigned* const diff = reinterpret_cast<unsigned*>(sm.clip);
    unsigned long long* const rowk = reinterpret_cast<unsigned long long*>(diff + TILE_H * DIFF_PITCH);   // [row][min, max]
    bool used_diff = false;
#endif
    if (skip_bg && q_beg < q_end) {
#if B2R_ROWDIFF
        for (int i = threadIdx.x; i < TILE_H * DIFF_PITCH; i += RASTER_THREADS) diff[i] = 0x80008000u;
#endif
        for (int i = threadIdx.x; i < TILE_PX; i += RASTER_THREADS) {   // one tile row per warp and iteration
            const unsigned long long k = sm.z[i];
            const unsigned long long rmin = warp_min_u64(k != z_init ? k : ~0ull), rmax = warp_max_u64(k != z_init ? k : 0ull);
#if B2R_ROWDIFF
            if (lane == 0) { rowk[2 * (i >> 5)] = rmin; rowk[2 * (i >> 5) + 1] = rmax; }
#endif
            kb_min = min(kb_min, rmin); kb_max = max(kb_max, rmax);
        }
        if (lane == 0) { sm.red_min[wid] = kb_min; sm.red_max[wid] = kb_max; }
        __syncthreads();
#pragma unroll
        for (int w = 0; w < RASTER_WARPS; ++w) { kb_min = min(kb_min, sm.red_min[w]); kb_max = max(kb_max, sm.red_max[w]); }
    }
    const bool any_cov = kb_min <= kb_max;
    int uniform = 0;
    if ((!skip_bg || any_cov) && !(B2R_SKIP & 2)) {
        const int* quad_list = B.quad_list + (size_t)view * B.quad_cap;
        const QuadRec* vquads = quads + (size_t)view * quad_stride;
        // SPLIT: this CTA's contiguous part of the pair list
        const int qs_beg = SPLIT ? q_beg + (int)(((long long)(q_end - q_beg) * part) / n_parts) : q_beg;
        const int qs_end = SPLIT ? q_beg + (int)(((long long)(q_end - q_beg) * (part + 1)) / n_parts) : q_end;
        const int n_pairs = qs_end - qs_beg;
        for (;;) {
#if B2R_CLASSIFY32
            // self-scheduling: a share of the remaining pairs per grab, at most 32 -- ONE LANE PER PAIR classifies it (the
            // four corners of its rectangle one after the other: independent chains of two float64 divisions each), so a
            // grab is one round of dependent loads (list entry -> quad record) for up to 32 pairs instead of eight
            int t0 = 0, grab = 0;
            if (lane == 0) {
                const int seen = sm.next_quad;  // a stale value only changes the grab size
                grab = max(1, min(32, (n_pairs - seen) / RASTER_WARPS));
                t0 = qs_beg + atomicAdd(&sm.next_quad, grab);
            }
            t0 = __shfl_sync(0xffffffffu, t0, 0);
            grab = __shfl_sync(0xffffffffu, grab, 0);
            if (t0 >= qs_end) break;
            const int t_hi = min(t0 + grab, qs_end);
            const int tg = t0 + lane;
            int g_entry = 0, g_state = 0;  // 0 skip, 1 process, 2 process and every covered pixel passes the z test
            if (tg < t_hi) {
                g_entry = quad_list[tg];
                const QuadRec& G = vquads[g_entry & (QUAD_FULL_BIT - 1)];
                const int gx0 = max((int)G.bx0, X0), gx1 = min((int)G.bx1, X1) - 1;
                const int gy0 = max((int)G.by0, Yb0), gy1 = min((int)G.by1, Y1) - 1;
                if (gx0 <= gx1 && gy0 <= gy1) {
                    g_state = 1;
                    if (skip_bg) {
                        unsigned long long kmin = ~0ull, kmax = 0ull;
                        int code = 0;
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const int cx = (c & 1) ? gx1 : gx0, cy = (c & 2) ? gy1 : gy0;
                            const double z = -(G.nx * (double)cx + G.ny * (double)cy + G.D) / G.nz;
                            const double den = V.zl_sum - z * V.zl_diff;
                            const double gz = V.zl_num / den;
                            const int cc = (gz == gz) ? (den > 0 ? 1 : (den < 0 ? 2 : 3)) : 3;
                            code = (c == 0 || code == cc) ? cc : 3;
                            const unsigned long long k = zkey(gz);
                            kmin = min(kmin, k); kmax = max(kmax, k);
                        }
                        if (code != 3) {
                            if (rh ? (kmin > kb_max) : (kmax < kb_min)) g_state = 0;
                            else if (rh ? (kmax <= kb_min) : (kmin >= kb_max)) g_state = 2;
                        }
                    }
                }
                B2R_STAT(0, 1); if (g_state == 0) B2R_STAT(1, 1); if (g_state == 2) B2R_STAT(6, 1);
            }
            unsigned todo = __ballot_sync(0xffffffffu, g_state != 0);
#else
            // guided self-scheduling: eight pairs per grab while the list is long, fewer towards its end, so the
            // warps reach the closing barrier together
            int t0 = 0, grab = 0;
            if (lane == 0) {
                const int seen = sm.next_quad;  // a stale value only changes the grab size
                grab = max(1, min(8, (n_pairs - seen) / (2 * RASTER_WARPS)));
                t0 = qs_beg + atomicAdd(&sm.next_quad, grab);
            }
            t0 = __shfl_sync(0xffffffffu, t0, 0);
            grab = __shfl_sync(0xffffffffu, grab, 0);
            if (t0 >= qs_end) break;
            const int t_hi = min(t0 + grab, qs_end);
            const int tg = t0 + (lane >> 2);
            int g_entry = 0, g_state = 0;  // 0 skip, 1 process, 2 process and every covered pixel passes the z test
            double g_z = 0.0;
            int g_code = 3;
            if (tg < t_hi) {
                g_entry = quad_list[tg];
                const QuadRec& G = vquads[g_entry & (QUAD_FULL_BIT - 1)];
                const int gx0 = max((int)G.bx0, X0), gx1 = min((int)G.bx1, X1) - 1;
                const int gy0 = max((int)G.by0, Yb0), gy1 = min((int)G.by1, Y1) - 1;
                if (gx0 <= gx1 && gy0 <= gy1) {
                    g_state = 1;
                    if (skip_bg) {
                        const int cx = (lane & 1) ? gx1 : gx0, cy = (lane & 2) ? gy1 : gy0;
                        const double z = -(G.nx * (double)cx + G.ny * (double)cy + G.D) / G.nz;
                        const double den = V.zl_sum - z * V.zl_diff;
                        g_z = V.zl_num / den;
                        g_code = (g_z == g_z) ? (den > 0 ? 1 : (den < 0 ? 2 : 3)) : 3;
                    }
                }
            }
            if (skip_bg) {
                unsigned long long kmin = zkey(g_z), kmax = kmin;
#pragma unroll
                for (int o = 1; o <= 2; o <<= 1) {
                    const int other = __shfl_xor_sync(0xffffffffu, g_code, o);
                    g_code = (g_code == other) ? g_code : 3;
                    kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
                    kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
                }
                if (g_state && g_code != 3) {
                    if (rh ? (kmin > kb_max) : (kmax < kb_min)) g_state = 0;
                    else if (rh ? (kmax <= kb_min) : (kmin >= kb_max)) g_state = 2;
                }
            }
            if ((lane & 3) == 0 && tg < t_hi) { B2R_STAT(0, 1); if (g_state == 0) B2R_STAT(1, 1); if (g_state == 2) B2R_STAT(6, 1); }
            unsigned todo = __ballot_sync(0xffffffffu, (lane & 3) == 0 && g_state != 0);
#endif
            while (todo) {
                const int src = __ffs(todo) - 1;
                todo &= todo - 1;
                const int entry = __shfl_sync(0xffffffffu, g_entry, src);
                const bool all_pass = __shfl_sync(0xffffffffu, g_state, src) == 2;
                const bool full = (entry & QUAD_FULL_BIT) != 0;
                const QuadRec& R = vquads[entry & (QUAD_FULL_BIT - 1)];
                const bool front = R.front != 0;
                if (all_pass && full) {
                    uniform += front ? 1 : -1;
                    if (lane == 0) B2R_STAT(4, 1);
                    continue;
                }
                const int rx0 = max((int)R.bx0, X0), rx1 = min((int)R.bx1, X1) - 1;
                const int ry0 = max((int)R.by0, Yb0), ry1 = min((int)R.by1, Y1) - 1;
                // exact span of row py = Y0 + lane: every edge function is monotone in px
                const int py = Y0 + lane;
                int lo = rx0, hi = rx1;
                if (py < ry0 || py > ry1) hi = lo - 1;
                const int nv = full ? 0 : R.n;
                if (full && lane == 0) B2R_STAT(5, 1);
                // Lane e < nv owns edge e: its vectors, and whether it constrains any row of the rectangle at all -- an edge
                // whose worst corner of the rectangle is already inside does not (most edges of a sliver crossing the
                // tile).  One lane per edge decides that; before, every lane decided it for every edge.
                double own_x = 0, own_y = 0, own_ex = 0, own_ey = 0;
                bool constrains = false;
                if (lane < nv) {
                    const int j = (lane + 1 == nv) ? 0 : lane + 1;
                    own_x = R.x[lane]; own_y = R.y[lane];
                    own_ex = R.x[j] - own_x; own_ey = R.y[j] - own_y;
                    const double fworst = front ? edge_fn(own_ey >= 0 ? rx0 : rx1, own_ex <= 0 ? ry0 : ry1, own_x, own_y, own_ex, own_ey)
                                                : edge_fn(own_ey >= 0 ? rx1 : rx0, own_ex <= 0 ? ry1 : ry0, own_x, own_y, own_ex, own_ey);
                    constrains = !(front ? (fworst > 0) : (fworst < 0));
                }
                unsigned edges_left = __ballot_sync(0xffffffffu, constrains);
                if (B2R_SKIP & 8) edges_left = 0;
                while (edges_left) {
                    const int e = __ffs(edges_left) - 1;
                    edges_left &= edges_left - 1;
                    const double xi = __shfl_sync(0xffffffffu, own_x, e), yi = __shfl_sync(0xffffffffu, own_y, e);
                    const double ex = __shfl_sync(0xffffffffu, own_ex, e), ey = __shfl_sync(0xffffffffu, own_ey, e);
                    if (lo > hi) continue;
                    const double c = ((double)py - yi) * ex;
                    auto pred = [&](int px) {  // front ? f > 0 : f < 0 with f = (px - xi)*ey - c  (triangular.py:305-311)
                        const double f = ((double)px - xi) * ey - c;
                        return front ? (f > 0) : (f < 0);
                    };
                    const bool up = front ? (ey > 0) : (ey < 0);  // the true set is upward closed in px
                    if (ey == 0 || !(ey == ey)) {
                        if (!pred(lo)) hi = lo - 1;
                        continue;
                    }
                    // where f changes sign, estimated in float32 (the walk below finds the exact cut whatever the estimate)
                    const float est = (float)xi + __fdividef((float)c, (float)ey);
                    if (up) {
                        if (!pred(hi)) { hi = lo - 1; continue; }
                        int k = est >= (float)hi ? hi : (est <= (float)lo ? lo : (int)ceilf(est));
                        if (!(est == est)) k = lo;
                        while (k > lo && pred(k - 1)) --k;
                        while (!pred(k)) ++k;
                        lo = k;
                    } else {
                        if (!pred(lo)) { hi = lo - 1; continue; }
                        int k = est >= (float)hi ? hi : (est <= (float)lo ? lo : (int)floorf(est));
                        if (!(est == est)) k = hi;
                        while (k < hi && pred(k + 1)) ++k;
                        while (!pred(k)) --k;
                        hi = k;
                    }
                }
                const int delta = front ? 1 : -1;
                if (B2R_SKIP & 16) continue;
#if B2R_ROWDIFF
                if (skip_bg) {
                    // row-level depth classification (lane = row): the quad depth, as the reference rounds it, is monotone
                    // along a row, so its values at the two ends of the span bound it over the span; against the range of
                    // the row's covered z-buffer entries the whole span fails (dropped), passes (difference array) or
                    // stays undecided (per-pixel items below).  `all_pass` pairs pass in every row.
                    bool row_pass = all_pass;
                    if (hi >= lo && !all_pass) {
                        const unsigned long long rmin = rowk[2 * lane], rmax = rowk[2 * lane + 1];
                        if (rmin > rmax) hi = lo - 1;   // no covered pixel in this row
                        else {
                            const double t = R.ny * (double)py;
                            const double z0 = -((R.nx * (double)lo + t) + R.D) / R.nz, z1 = -((R.nx * (double)hi + t) + R.D) / R.nz;
                            const double den0 = V.zl_sum - z0 * V.zl_diff, den1 = V.zl_sum - z1 * V.zl_diff;
                            const double q0 = V.zl_num / den0, q1 = V.zl_num / den1;
                            if (q0 == q0 && q1 == q1 && ((den0 > 0 && den1 > 0) || (den0 < 0 && den1 < 0))) {
                                const unsigned long long k0 = zkey(q0), k1 = zkey(q1);
                                const unsigned long long kmin = min(k0, k1), kmax = max(k0, k1);
                                if (rh ? (kmin > rmax) : (kmax < rmin)) hi = lo - 1;          // every covered pixel fails
                                else if (rh ? (kmax <= rmin) : (kmin >= rmax)) row_pass = true;  // every covered pixel passes
                            }
                        }
                    }
                    if (row_pass && hi >= lo) {
                        const int a = lo - X0, b = hi - X0 + 1;
                        unsigned* const drow = diff + lane * DIFF_PITCH;
                        atomicAdd(drow + (a >> 1), (unsigned)delta << (16 * (a & 1)));
                        if (b < TILE_W) atomicAdd(drow + (b >> 1), (unsigned)(-delta) << (16 * (b & 1)));
                        used_diff = true;
                        hi = lo - 1;
                    }
                }
#endif
                // The spans of the 32 rows are flattened into one dense pixel list and dealt to the lanes, so every lane
                // works whatever the shape of the quad: pixel k lies in the row r with start[r] <= k < start[r+1]
                // (inclusive scan over the lanes, then a 5-step bisection through shuffles).  Two pixels per lane and
                // iteration, written as straight-line code: their depth evaluations (two dependent float64 divisions
                // each) are independent and overlap in the pipeline.
                if (!__any_sync(0xffffffffu, hi >= lo)) continue;   // every row was dropped or went to the difference array
                const int len = max(hi - lo + 1, 0);
                int incl = len;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
                const int total = __shfl_sync(0xffffffffu, incl, 31);
                if (lane == 0) { B2R_STAT(2, 1); B2R_STAT(3, total); }
                // row table: entry k of the warp's table names the row of item k (written by the lane that owns the row);
                // an item is then one shared-memory byte and one shuffle away from its pixel (before: a 5-step bisection
                // through shuffles per item)
                const int off = lo - X0 - (incl - len);   // lx = k + off for the items k of this lane's row
                __syncwarp();                             // the previous pair's table is fully consumed
                for (int j = incl - len; j < incl; ++j) rowtab[j] = (unsigned char)lane;
                __syncwarp();
                auto locate = [&](int k, bool valid, int& lx, int& r) {
                    r = valid ? (int)rowtab[k] : 0;
                    lx = k + __shfl_sync(0xffffffffu, off, r);
                };
                auto quad_depth = [&](int px, int qy) {  // z = -(nx*px + ny*py + D)/nz, linearised (triangular.py:352-354)
                    const double z = -(R.nx * (double)px + R.ny * (double)qy + R.D) / R.nz;
                    return V.zl_num / (V.zl_sum - z * V.zl_diff);
                };
                // what is left for the per-pixel path after the row classification is a tenth of the items: one pixel per
                // lane and iteration keeps the loop (and the kernel's instruction footprint) small
                for (int k = lane; k < ((total + 31) & ~31); k += 32) {
                    int lx, r;
                    const bool valid = k < total;
                    locate(k, valid, lx, r);
                    const int p = valid ? tpix(lx, r) : 0;
                    const unsigned long long kb = sm.z[p];
                    bool hit = valid && !(skip_bg && kb == z_init);
                    if (!all_pass) {
                        const double zq = quad_depth(X0 + lx, Y0 + r);
                        const unsigned long long kz = zkey(zq);
                        hit = hit && zq == zq && (rh ? (kb >= kz) : (kb <= kz));
                    }
                    if (hit) atomicAdd(&sm.st[p], delta);
                }
            }  // survivors of this grab
        }
    }
    if (skip_bg && lane == 0 && uniform) atomicAdd(&sm.uniform, uniform);
#if B2R_ROWDIFF
    if (__any_sync(0xffffffffu, used_diff) && lane == 0) sm.use_diff = 1;
#endif
    __syncthreads();
#if B2R_ROWDIFF
    if (sm.use_diff) {   // difference arrays -> counts: a warp per row, inclusive scan over the 32 pixels
        for (int row = wid; row < TILE_H; row += RASTER_WARPS) {
            const unsigned w = diff[row * DIFF_PITCH + (lane >> 1)];
            int v = (int)((w >> (16 * (lane & 1))) & 0xffffu) - 0x8000;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += u; }
            if (v) sm.st[tpix(lane, row)] += v;
        }
        __syncthreads();
    }
#endif
    if (skip_bg && sm.uniform) {
        const int uadd = sm.uniform;
        for (int i = threadIdx.x; i < TILE_PX; i += RASTER_THREADS) if (sm.z[i] != z_init) sm.st[i] += uadd;
        __syncthreads();
    }

    if (SPLIT && n_parts > 1 && !solo) {
        // partial stencil counts -> global memory; the last part of the tile to arrive sums them and goes on alone
        const size_t slot = (size_t)view * n_tiles + tile;   // view = position in the batch: concurrent sub-chunks use disjoint slots
        int* const mine = O.split_st + (slot * n_parts + part) * TILE_PX;
        for (int i = threadIdx.x; i < TILE_PX; i += RASTER_THREADS) mine[i] = sm.st[i];
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) sm.ticket = atomicAdd(O.split_ticket + slot, 1);
        __syncthreads();
        if (sm.ticket != n_parts - 1) return;
        __threadfence();
        for (int q = 0; q < n_parts; ++q) {
            if (q == part) continue;
            const int* other = O.split_st + (slot * n_parts + q) * TILE_PX;
            for (int i = threadIdx.x; i < TILE_PX; i += RASTER_THREADS) sm.st[i] += __ldcg(other + i);
        }
        __syncthreads();
    }
    if (sm.need_full) {
        if (threadIdx.x == 0) B2R_STAT(7, 1);
        for (int i = threadIdx.x; i < TILE_PX; i += RASTER_THREADS) sm.id[i] = -1;
        tile_tris<3>(sm, S.pos, S.face_vf, V, vtris, tri_list, t_beg, t_end, X0, Y0, X1, Yb0, Y1, rh, status_view);
        __syncthreads();
    } else if (status_view) {
        // Pass-3 status (core.py:624-636) without the full pass: no exact tie was seen in this tile, so the faces that pass
        // the z-test here are exactly the verified winners, and a face is "rendered" iff it wins a pixel whose stencil
        // count is 0 (every face that reached the lists covers a pixel: k_tri_setup set the covered bit).  A face's
        // byte is read first: the atomic is issued once per face and tile, not once per pixel.
        for (int i = threadIdx.x; i < TILE_PX; i += RASTER_THREADS) {
            const int f = sm.id[i];
            if (f < 0 || sm.st[i] != 0) continue;
            uint8_t* sp = status_view + f;
            if (*(volatile uint8_t*)sp & 4) continue;
            unsigned* wp = (unsigned*)((uintptr_t)sp & ~(uintptr_t)3);
            atomicOr(wp, (2u | 4u) << (8 * ((uintptr_t)sp & 3)));
        }
    }

    // ---- output: a warp per tile row ----
    const bool dbg_planes = O.winner || O.stencil || O.z;
    for (int row = wid; row < TILE_H; row += RASTER_WARPS) {
        const int py = Y0 + row, px = X0 + lane;
        if (py < Yb0 || py >= Y1) continue;  // whole warp
        const int p = tpix(lane, row);
        const int face = sm.id[p], stc = sm.st[p];
        const size_t g = plane + (size_t)py * Fr.W + px;
        if (dbg_planes && px < X1) {
            if (O.winner) O.winner[g] = face;
            if (O.stencil) O.stencil[g] = (short)stc;
            if (O.z) O.z[g] = zkey_decode(sm.z[p]);
        }
        if (!FUSED) {   // one packed word per pixel for k_shade_packed
            if (px < X1) O.packed[g] = face < 0 ? PACKED_NONE : ((unsigned)face | (stc == 0 ? PACKED_LIT : 0u));
            continue;
        }
        // shading straight from the shared planes (triangular.py:135-171, core.py:640)
        float c[3] = {0.f, 0.f, 0.f};
        bool const_bg = false;
        if (px < X1) {
            if (face >= 0) {
                if (!shade_face_pixel(S, V, Fr.light, vtris[face], face, px, py, stc == 0, c)) *Fr.err_flag = 1;
            } else if (Fr.bg_mode == B2R_BG_CUBEMAP) {
                skybox_pixel(S, V, Fr.sky_size, px, py, c);
            } else {
                c[0] = Fr.background[0]; c[1] = Fr.background[1]; c[2] = Fr.background[2];
                const_bg = true;
            }
            if (O.f32) { float* o = O.f32 + g * 3; o[0] = c[0]; o[1] = c[1]; o[2] = c[2]; }
        }
        const unsigned packed = const_bg ? __ldg(O.bg_packed) : tonemap_pack(c);
        store_row(Fr, O.rgb, view, X0, py, packed, px < X1, lane);
    }
}

// Shading pass of the unfused path: one CTA of 128 threads per 32x32 tile (a warp per row), reading the packed winner
// word k_tile<false> left behind.  Tiles whose lists are empty were never written: they are recognised here.
// MODE: SHADE_F32 (production) general_shading with float32 lighting, SHADE_F64 the all-float64 form (B2R_SHADE_F64=1,
// and whenever the float frame is handed out), SHADE_ALT one of the alternative shading functions
// (Fr.shading != B2R_SHADE_GENERAL); separate instantiations so that the production kernel carries none of the others'
// code or stack.
enum : int { SHADE_F64 = 0, SHADE_ALT = 1, SHADE_F32 = 2 };
template <int MODE>
__global__ void __launch_bounds__(RASTER_THREADS, B2R_SHADE_MINB)
k_shade_packed(SceneDev S, const ViewDev* __restrict__ views, FrameDev Fr, const TriRec* __restrict__ tris, BinDev B,
               TileOut O, int view0, int n_sub) {
    const int view = (int)blockIdx.x + view0;
    const ViewDev& V = views[view];
    const int n_tiles = Fr.tiles_x * Fr.tiles_y;
    const int tile = B.order[(size_t)view * n_tiles + blockIdx.y];
    const int tx = tile % Fr.tiles_x, ty = tile / Fr.tiles_x + Fr.tile_row0;
    const int X0 = tx * TILE_W, Y0 = ty * TILE_H;
    const int X1 = min(X0 + TILE_W, Fr.W), Y1 = min(min(Y0 + TILE_H, Fr.H), Fr.row_end);
    const int Yb0 = max(Y0, Fr.row_begin);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int* tri_off = B.tri_off + (size_t)view * (n_tiles + 1);
    const int* quad_off = B.quad_off + (size_t)view * (n_tiles + 1);
    const bool lists_ok = (B.overflow[view * 2] | B.overflow[view * 2 + 1]) == 0;
    const bool empty = !lists_ok || (tri_off[tile] == tri_off[tile + 1] &&
                                     (quad_off[tile] == quad_off[tile + 1] || !Fr.full_stencil));
    const size_t plane = (size_t)view * Fr.H * Fr.W;
    const TriRec* vtris = tris + (size_t)view * Fr.n_faces;
    const unsigned bg = __ldg(O.bg_packed);
    // a row segment that is constant background everywhere (all of an empty tile, most rows of a partly covered one):
    // lane j < 24 stores bytes 4j .. 4j+3 of the repeating R,G,B pattern, no shuffles, no tonemap
    const bool fast_bg = Fr.bg_mode != B2R_BG_CUBEMAP && !O.f32 && X0 + 32 <= Fr.W && (((size_t)Fr.W * 3) & 3) == 0;
    unsigned bg_word = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) bg_word |= ((bg >> (8 * ((4 * lane + i) % 3))) & 0xffu) << (8 * i);
    for (int row = wid; row < TILE_H; row += RASTER_WARPS) {
        const int py = Y0 + row, px = X0 + lane;
        if (py < Yb0 || py >= Y1) continue;  // whole warp
        const size_t g = plane + (size_t)py * Fr.W + px;
        const unsigned w = (empty || px >= X1) ? PACKED_NONE : O.packed[g];
        if (fast_bg && __all_sync(0xffffffffu, w == PACKED_NONE)) {
            if (lane < 24)
                reinterpret_cast<unsigned*>(O.rgb + ((size_t)view * Fr.H + (size_t)(Fr.H - 1 - py)) * Fr.W * 3 + (size_t)X0 * 3)[lane] = bg_word;
            continue;
        }
        float c[3] = {0.f, 0.f, 0.f};
        bool const_bg = false;
        if (px < X1) {
            if (w != PACKED_NONE) {
                const int face = (int)(w & ~PACKED_LIT);
                if (MODE == SHADE_ALT) shade_alt_pixel(S, V, Fr.light, vtris[face], face, px, py, Fr.shading, c);
                else if (MODE == SHADE_F32) {
                    if (!shade_face_pixel_f32(S, V, Fr.lightf, Fr.light.type, vtris[face], face, px, py, (w & PACKED_LIT) != 0, c)) *Fr.err_flag = 1;
                } else if (!shade_face_pixel(S, V, Fr.light, vtris[face], face, px, py, (w & PACKED_LIT) != 0, c)) *Fr.err_flag = 1;
            } else if (Fr.bg_mode == B2R_BG_CUBEMAP) {
                skybox_pixel(S, V, Fr.sky_size, px, py, c);
            } else {
                c[0] = Fr.background[0]; c[1] = Fr.background[1]; c[2] = Fr.background[2];
                const_bg = true;
            }
            if (O.f32) { float* o = O.f32 + g * 3; o[0] = c[0]; o[1] = c[1]; o[2] = c[2]; }
        }
        const unsigned packed = const_bg ? bg : (MODE == SHADE_ALT ? tonemap_pack_wide(c) : tonemap_pack(c));
        store_row(Fr, O.rgb, view, X0, py, packed, px < X1, lane);
    }
}

// pass-3 status of every face (core.py:624-636): pending faces carry coverage / z / lit bits ORed in by k_tile
__global__ void k_status_resolve(uint8_t* __restrict__ status, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t s = status[i];
    if ((s & 0xE0) != 0xE0) return;
    status[i] = (s & 4) ? B2R_FACE_RENDERED : ((s & 1) ? B2R_FACE_EMPTY_Z : B2R_FACE_CLIPPED);
}

// =====================================================================================================================
// Multi-GPU output window: sparse push of finished frames into the assembling rank's buffer over NVLink.
// One 32x32 tile at a time: the tile's 32 row segments (96 bytes each) are read from the local frames with 16-byte
// loads; if every pixel of the tile has the colour of its first pixel AND the destination tile is known to hold exactly
// that (state: what this rank pushed into the same window block the last time), nothing is stored -- the constant
// background of a frame (about half of the headline scene's tiles) never crosses the link again.  Otherwise the tile is
// stored with 16-byte peer stores and the state updated.  Purely content based, hence exact whatever was rendered.
// Why: seven peers pushing whole frames deliver 7 x 398 MB per 64-frame step into ONE 900 GB/s NVLink port; with the
// renderers at 3.4 ms per step that port (748 GB/s measured) bounded the 8-GPU weak-scaling efficiency at 0.92.
// Persistent CTAs (a low-priority stream, under the next step's render), 192 threads = one 16-byte piece of the tile each.
// Requires W % 32 == 0 (rows and tile segments are then 16-byte aligned); the host wrapper falls back to a plain copy.
// =====================================================================================================================
constexpr int PUSH_THREADS = 128;   // four warps, a tile each: 32 lanes x 6 pieces of 16 bytes = the tile's 32 rows of 96 bytes
__global__ void __launch_bounds__(PUSH_THREADS, 8) k_window_push(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int n_views,
                                                              int H, int W, unsigned* __restrict__ state) {
    const int tiles_x = W / TILE_W, tiles_y = (H + TILE_H - 1) / TILE_H;
    const int n_tiles = tiles_x * tiles_y, total = n_views * n_tiles;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * PUSH_THREADS + threadIdx.x) >> 5, n_warps = (gridDim.x * PUSH_THREADS) >> 5;
    for (int item = warp; item < total; item += n_warps) {
        const int view = item / n_tiles, tile = item - view * n_tiles;
        const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
        const size_t tile_off = ((size_t)view * H + (size_t)ty * TILE_H) * W * 3 + (size_t)tx * (TILE_W * 3);
        const int rows = min(TILE_H, H - ty * TILE_H);
        // six independent 16-byte loads per lane: piece q = lane + 32 j of the tile's 192 (row q / 6, 16-byte column q % 6)
        uint4 v[6];
        auto piece_off = [&](int j) {
            const int q = lane + 32 * j, row = q / 6, col = q - row * 6;
            return tile_off + (size_t)min(row, rows - 1) * W * 3 + (size_t)col * 16;
        };
#pragma unroll
        for (int j = 0; j < 6; ++j) v[j] = __ldg(reinterpret_cast<const uint4*>(src + piece_off(j)));
        const unsigned c = __shfl_sync(0xffffffffu, v[0].x, 0) & 0x00ffffffu;   // colour of the tile's first pixel
        // A row of a tile filled with colour c = (R, G, B) repeats three words: RGBR GBRG BRGB.  Word k of the piece at
        // 16-byte column col is word (4 col + k) % 3 = (col + k) % 3 of that pattern.
        const unsigned c0 = c & 0xffu, c1 = (c >> 8) & 0xffu, c2 = (c >> 16) & 0xffu;
        const unsigned pat[3] = {c0 | c1 << 8 | c2 << 16 | c0 << 24, c1 | c2 << 8 | c0 << 16 | c1 << 24, c2 | c0 << 8 | c1 << 16 | c2 << 24};
        const int r0 = lane % 3;   // col % 3 of piece j: (lane + 32 j) % 3 = (r0 + 2 j) % 3
        bool same = true;
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            const int q = lane + 32 * j;
            const int r = (r0 + 2 * j) % 3;
            const unsigned w0 = r == 0 ? pat[0] : (r == 1 ? pat[1] : pat[2]);
            const unsigned w1 = r == 0 ? pat[1] : (r == 1 ? pat[2] : pat[0]);
            const unsigned w2 = r == 0 ? pat[2] : (r == 1 ? pat[0] : pat[1]);
            same = same && (q / 6 >= rows || (v[j].x == w0 && v[j].y == w1 && v[j].z == w2 && v[j].w == w0));
        }
        const bool uniform = __all_sync(0xffffffffu, same);
        const unsigned tag = 0x01000000u | c;   // "the destination tile holds nothing but colour c"
        const unsigned held = state[item];
        if (uniform && held == tag) continue;
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            const int q = lane + 32 * j;
            if (q / 6 < rows) *reinterpret_cast<uint4*>(dst + piece_off(j)) = v[j];
        }
        if (lane == 0) state[item] = uniform ? tag : 0u;
    }
}

}  // namespace b2r
