// b2r_api.cu -- host side of the C ABI declared in include/b2r.h (scene upload, per-frame orchestration).
//
// One stream, no host synchronisation inside a frame: every stage is a kernel launch on `g.stream`; sizes that
// are only known on the device (silhouette length, tile-list totals) stay on the device (grid-stride kernels read
// them), with capacity overflow reported through a flag that is read back with the frame.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "b2r_kernels.cuh"

using namespace b2r;

namespace {

thread_local std::string t_error;

constexpr int MAX_FLAG_VIEWS = 4096;
constexpr int N_TICKET_SLOTS = 4;
constexpr int FLAG_REGIONS = N_TICKET_SLOTS + 2;   // + one for synchronous host renders, one scratch region (device-out)
constexpr int STICKY_SLOTS = 1024;
constexpr int STICKY_INTS = 4;                     // need_tri, need_quad, texture-index error, pad
constexpr int REGION_INTS = 2 * MAX_FLAG_VIEWS + 16;  // 2 ints per view + [2*MAX_FLAG_VIEWS] = texture-index error word

// Everything the library owns on ONE CUDA device: streams, events, pinned staging.  There is no process-global
// render state: b2r_init(device) creates (or selects) the context of that device, a scene is bound to the context it
// was created on, and two scenes on two devices can live -- and render from two threads -- in one process.  Entry
// points lock the context they act on.
struct Context {
    std::recursive_mutex mu;
    bool ready = false;
    int device = -1;
    int sm_count = 148;
    size_t scratch_budget = (size_t)6 << 30;   // per-call scratch of a batch: a sixth of the device memory, at most 32 GiB (B200: 30 GiB)
    cudaStream_t stream = nullptr;
    long long launches = 0;
    bool timing = false;
    int n_stage = 0;
    const char* stage_name[B2R_MAX_STAGES];
    cudaEvent_t stage_ev[B2R_MAX_STAGES + 1] = {};
    int* pinned_flags = nullptr;  // overflow read-back, FLAG_REGIONS x (2 ints per view)
    int* sticky = nullptr;        // pinned, STICKY_SLOTS x STICKY_INTS: "a tile list overflowed" per scene, survives until b2r_sync
    std::vector<int> sticky_free;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t chunk_done = nullptr;
    static constexpr int N_AUX = 3;   // sub-chunks of one batch render concurrently: the long tail of a raster launch
    cudaStream_t aux[N_AUX] = {};     // (a few very heavy tiles) is filled by the next sub-chunk's work
    cudaEvent_t aux_done[16] = {};
    cudaEvent_t aux_last[N_AUX] = {};  // last tile / shading work enqueued on each auxiliary stream
    cudaEvent_t setup_done = nullptr;
    cudaStream_t fork_stream = nullptr;  // k_tri_setup runs here, next to the facing -> silhouette -> quad chain
    cudaEvent_t tri_done = nullptr;
    int host_chunk = 2;           // views per sub-chunk when frames go to host memory (copy/compute overlap)
    int dev_chunk = 64;           // views per tile launch when frames stay on the device (a whole pipeline part: the cost-ordered grid has no tail to hide)
    int pipe_views = 1 << 20;     // B2R_PIPE: views per pipeline part (set-up + binning of part p+1 under the tile kernels of part p).
                                  // OFF by default: measured on the B200 (diablo, 64 views) 13.5 k frames/s unsplit against 12.6 k with
                                  // parts of 8 / 16 / 32 views -- the cost-ordered tile grid loses more to its shorter launches than
                                  // the 0.4 ms of set-up it hides
    int async_chunk = 8;          // views per sub-chunk of a host-asynchronous call (swept: tools/knob_sweep_e2e.sh)
    int aux_host = 2, aux_dev = 3;  // how many auxiliary streams the sub-chunks rotate over
    int init_tri_cap = 0, init_quad_cap = 0;  // B2R_TRI_CAP / B2R_QUAD_CAP: first per-view list capacities (tests of the grow path)
    int bin_blocks = 0, bin_share = 64;  // k_bin grid (0 = 2 per SM) and the most warps that share one quad
    int split_parts = 4;          // B2R_SPLIT: CTAs per tile in launches of one or two views (1 = off)
    bool clip_elide = true;       // B2R_CLIP_ELIDE=0: keep the per-pixel clip test on every (clip-flagged face, tile) pair (A/B)
    bool debug_skip_bg = false;   // B2R_DEBUG_SKIP_BG=1: debug stencil plane from the production (skip-background) stencil path
    bool shade_f32 = true;        // B2R_SHADE_F64=1: the all-float64 shading kernel instead of float32 lighting
    int fused = 0;                // B2R_FUSED: 0 = k_tile<false> + k_shade_packed (one packed word per pixel between them; production),
                                  // 1 = k_tile<true> (shading inside the tile kernel; measured slower, kept for A/B)
    // pinned staging ring for the per-view constants: a pageable source would make cudaMemcpyAsync synchronise the
    // stream, i.e. serialise the host with the previous chunk / previous asynchronous call
    struct Staging { ViewDev* host = nullptr; size_t cap = 0; cudaEvent_t done = nullptr; bool used = false; } stage[4];
    unsigned stage_next = 0;
    // host-asynchronous renders (out_on_device == 2) -- and only those -- take a ticket; ticket t lives in slot t % 4
    // until b2r_wait(t) or until the slot is needed again (it is then completed first, its outcome remembered)
    long long ticket = 0;
    struct TicketSlot { long long id = 0; bool open = false; int views = 0; struct b2r_scene* scene = nullptr;
                        cudaEvent_t copy = nullptr, compute = nullptr; } tk[N_TICKET_SLOTS];
    std::map<long long, std::string> late_fail;  // tickets completed at slot reuse that had failed
    std::vector<struct b2r_scene*> pending;      // scenes with device-resident renders since the last b2r_sync
};

constexpr int MAX_DEVICES = 64;
std::mutex g_table_mu;
Context* g_ctx[MAX_DEVICES] = {};
std::atomic<int> g_last_device{-1};
thread_local int t_device = -1;

// the context of the calling thread: the device it last passed to b2r_init, else the most recently initialised one
Context* current_ctx() {
    const int d = t_device >= 0 ? t_device : g_last_device.load();
    if (d < 0 || d >= MAX_DEVICES) return nullptr;
    Context* c = g_ctx[d];
    return (c && c->ready) ? c : nullptr;
}

int fail(const std::string& msg) {
    t_error = msg;
    return 1;
}
int fail_index() {  // the reference raises IndexError: a texture lookup below -size (core.py:138-143, 162-173)
    t_error = "texture lookup out of range (UV < -1): index out of bounds";
    return B2R_ERR_INDEX;
}
#define CK(call)                                                                                        \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess)                                                                          \
            return fail(std::string(#call) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + ":" +   \
                        std::to_string(__LINE__) + ")");                                                \
    } while (0)
// entry points without a scene argument act on the calling thread's context
#define CTX_OR_FAIL()                                                     \
    Context* cxp_ = current_ctx();                                        \
    if (!cxp_) return fail("b2r_init was not called");                    \
    Context& g = *cxp_;                                                   \
    std::lock_guard<std::recursive_mutex> lock_(g.mu)

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    cudaError_t reserve(size_t count) {
        if (count <= n) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
        cudaError_t e = cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T) + 16);
        if (e == cudaSuccess) n = count;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    size_t bytes() const { return n * sizeof(T); }
};

void stage_mark(Context& g, const char* name) {
    if (!g.timing || g.n_stage >= B2R_MAX_STAGES) return;
    g.stage_name[g.n_stage] = name;
    cudaEventRecord(g.stage_ev[g.n_stage + 1], g.stream);
    ++g.n_stage;
}

}  // namespace

struct b2r_scene {
    Context* cx = nullptr;   // the device context this scene lives on
    int sticky_slot = -1;    // index into cx->sticky (overflow flags of device-resident renders)
    bool in_pending = false;
    // static geometry
    DevBuf<double4> pos;
    DevBuf<double2> uv;
    DevBuf<double> nrm;
    DevBuf<FaceStatic> faces;
    DevBuf<int4> face_vf;
    DevBuf<ShadeStatic> shade;
    DevBuf<ShadeLite> shade_lite;
    DevBuf<MaterialDev> mats;
    DevBuf<TextureDev> tex;
    std::vector<uchar4*> tex_data;
    DevBuf<uchar4> sky;
    int sky_size = 0;
    DevBuf<int2> edge_v;
    DevBuf<int> edge_ptr, edge_inc, edge_model;
    DevBuf<int8_t> sil_state;
    int n_faces = 0, n_edges = 0, n_models = 0;
    bool has_no_zwrite = false;  // some model has depth_test == False
    std::vector<int> model_faces;
    size_t static_bytes = 0;
    // per-render scratch (grown on demand)
    DevBuf<uint8_t> facing;
    DevBuf<SilEdge> sil;
    DevBuf<int> counters;  // [0] silhouette count, [1..n_models] per-model counts, [1+n_models] tonemapped background
    DevBuf<ViewDev> views;
    DevBuf<TriRec> tris;
    DevBuf<TriBox> boxes;     // (views, F) what binning needs of a face
    DevBuf<double4> vrec;     // (views, V) per-vertex screen x, y, z and 1/w (k_vertex)
    DevBuf<uint8_t> vinside;  // (views, V) vertex well inside both frusta
    int n_vertices = 0;
    DevBuf<int> coop_list;    // (views, F) faces queued for k_tri_count; their counts sit behind the tile counters
    DevBuf<QuadRec> quads;
    DevBuf<int> tile_counts, tile_offs, tri_list, quad_list, overflow, tile_order;
    DevBuf<int2> pair_list;  // (quad, tile) pairs between the two binning passes
    DevBuf<int> winner;
    DevBuf<unsigned> packed;  // winner | lit << 31 per pixel (B2R_FUSED=2)
    DevBuf<int> split_st, split_ticket;  // SPLIT launches of the tile kernel: partial stencil counts, tickets
    DevBuf<short> stencil;
    DevBuf<double> zplane;
    DevBuf<float> frame_f32;
    DevBuf<uint8_t> status;
    DevBuf<uint8_t> rgb[2];  // device staging of host-bound frames, alternating so that copies of call t overlap call t+1
    cudaEvent_t rgb_copied[2] = {nullptr, nullptr};  // the frames of the host render that last used rgb[i] have left it
    bool rgb_used[2] = {false, false};
    unsigned host_seq = 0;   // host-bound renders of this scene so far (selects rgb[host_seq & 1])
    int tri_cap = 0, quad_cap = 0;  // per-view capacity of the tile lists (grown after an overflow)
    int sil_cap = 0;                // capacity of the silhouette / per-view quad records (edges that can be extruded at once)

    SceneDev dev() const {
        SceneDev S;
        S.pos = pos.p; S.uv = uv.p; S.nrm = nrm.p; S.faces = faces.p; S.face_vf = face_vf.p; S.shade = shade.p; S.shade_lite = shade_lite.p; S.mats = mats.p; S.tex = tex.p; S.sky = sky.p;
        S.edge_v = edge_v.p; S.edge_ptr = edge_ptr.p; S.edge_inc = edge_inc.p;
        S.n_faces = n_faces; S.n_edges = n_edges;
        return S;
    }
};

static void b2r_scene_grow_lists(struct b2r_scene* sc, int need_tri, int need_quad, int need_sil = 0);
static int b2r_scene_check_sticky(struct b2r_scene* sc);

extern "C" {

int b2r_abi_version(void) { return B2R_ABI_VERSION; }
const char* b2r_last_error(void) { return t_error.c_str(); }

int b2r_init(int device) {
    if (device < 0 || device >= MAX_DEVICES) return fail("device index out of range");
    std::lock_guard<std::mutex> table_lock(g_table_mu);
    if (g_ctx[device] && g_ctx[device]->ready) { t_device = device; g_last_device = device; return 0; }
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(std::string("no CUDA device available (") + cudaGetErrorString(e) +
                    "); this library has no CPU fallback");
    if (device >= count) return fail("device index out of range");
    CK(cudaSetDevice(device));
    if (!g_ctx[device]) g_ctx[device] = new Context();
    Context& g = *g_ctx[device];
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    g.sm_count = prop.multiProcessorCount;
    g.scratch_budget = std::max((size_t)2 << 30, std::min((size_t)32 << 30, (size_t)prop.totalGlobalMem / 6));
    if (const char* a = std::getenv("B2R_SCRATCH_GB")) g.scratch_budget = (size_t)std::max(1, std::atoi(a)) << 30;
    // the main stream carries the small, latency-bound set-up launches of the pipeline: highest priority, so that they
    // are not queued behind the tile kernels (auxiliary streams, default priority) they are meant to overlap
    int prio_lo = 0, prio_hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    CK(cudaStreamCreateWithPriority(&g.stream, cudaStreamNonBlocking, prio_hi));
    for (int i = 0; i <= B2R_MAX_STAGES; ++i) CK(cudaEventCreate(&g.stage_ev[i]));
    CK(cudaMallocHost(&g.pinned_flags, sizeof(int) * REGION_INTS * FLAG_REGIONS));
    std::memset(g.pinned_flags, 0, sizeof(int) * REGION_INTS * FLAG_REGIONS);
    CK(cudaMallocHost(&g.sticky, sizeof(int) * STICKY_INTS * STICKY_SLOTS));
    std::memset(g.sticky, 0, sizeof(int) * STICKY_INTS * STICKY_SLOTS);
    g.sticky_free.clear();
    for (int i = STICKY_SLOTS - 1; i >= 0; --i) g.sticky_free.push_back(i);
    for (int i = 0; i < N_TICKET_SLOTS; ++i) {
        CK(cudaEventCreateWithFlags(&g.tk[i].copy, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&g.tk[i].compute, cudaEventDisableTiming));
    }
    CK(cudaStreamCreateWithFlags(&g.copy_stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&g.chunk_done, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&g.setup_done, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&g.tri_done, cudaEventDisableTiming));
    CK(cudaStreamCreateWithPriority(&g.fork_stream, cudaStreamNonBlocking, prio_hi));
    for (int i = 0; i < Context::N_AUX; ++i) CK(cudaStreamCreateWithFlags(&g.aux[i], cudaStreamNonBlocking));
    for (int i = 0; i < 16; ++i) CK(cudaEventCreateWithFlags(&g.aux_done[i], cudaEventDisableTiming));
    for (int i = 0; i < Context::N_AUX; ++i) CK(cudaEventCreateWithFlags(&g.aux_last[i], cudaEventDisableTiming));
    if (const char* hc = std::getenv("B2R_HOST_CHUNK")) g.host_chunk = std::max(1, std::atoi(hc));
    if (const char* dc = std::getenv("B2R_DEV_CHUNK")) g.dev_chunk = std::max(1, std::atoi(dc));
    if (const char* ac = std::getenv("B2R_ASYNC_CHUNK")) g.async_chunk = std::max(1, std::atoi(ac));
    if (const char* a = std::getenv("B2R_AUX_HOST")) g.aux_host = std::min(Context::N_AUX, std::max(1, std::atoi(a)));
    if (const char* a = std::getenv("B2R_TRI_CAP")) g.init_tri_cap = std::max(1, std::atoi(a));
    if (const char* a = std::getenv("B2R_QUAD_CAP")) g.init_quad_cap = std::max(1, std::atoi(a));
    if (const char* a = std::getenv("B2R_BIN_BLOCKS")) g.bin_blocks = std::max(1, std::atoi(a));
    if (const char* a = std::getenv("B2R_BIN_SHARE")) g.bin_share = std::max(1, std::atoi(a));
    if (const char* a = std::getenv("B2R_PIPE")) g.pipe_views = std::max(1, std::atoi(a));
    if (const char* a = std::getenv("B2R_FUSED")) g.fused = std::atoi(a) != 0;
    if (const char* a = std::getenv("B2R_SPLIT")) g.split_parts = std::min(8, std::max(1, std::atoi(a)));
    if (const char* a = std::getenv("B2R_CLIP_ELIDE")) g.clip_elide = std::atoi(a) != 0;
    if (const char* a = std::getenv("B2R_DEBUG_SKIP_BG")) g.debug_skip_bg = std::atoi(a) != 0;
    if (const char* a = std::getenv("B2R_SHADE_F64")) g.shade_f32 = std::atoi(a) == 0;
    if (const char* a = std::getenv("B2R_AUX_DEV")) g.aux_dev = std::min(Context::N_AUX, std::max(1, std::atoi(a)));
    g.device = device;
    g.ready = true;
    g.launches = 0;
    t_device = device;
    g_last_device = device;
    return 0;
}

// Releases every context: all streams are drained first, then every resource b2r_init / b2r_render created is
// destroyed.  Scenes must be destroyed before (a scene outliving its context only frees its own allocations).
int b2r_shutdown(void) {
    std::lock_guard<std::mutex> table_lock(g_table_mu);
    for (int d = 0; d < MAX_DEVICES; ++d) {
        Context* c = g_ctx[d];
        if (!c) continue;
        {
            std::lock_guard<std::recursive_mutex> lock(c->mu);
            Context& g = *c;
            if (g.ready) {
                cudaSetDevice(g.device);
                cudaStreamSynchronize(g.stream); cudaStreamSynchronize(g.copy_stream); cudaStreamSynchronize(g.fork_stream);
                for (int i = 0; i < Context::N_AUX; ++i) cudaStreamSynchronize(g.aux[i]);
                cudaStreamDestroy(g.stream); cudaStreamDestroy(g.copy_stream); cudaStreamDestroy(g.fork_stream);
                for (int i = 0; i < Context::N_AUX; ++i) cudaStreamDestroy(g.aux[i]);
                cudaEventDestroy(g.chunk_done); cudaEventDestroy(g.setup_done); cudaEventDestroy(g.tri_done);
                for (int i = 0; i < 16; ++i) cudaEventDestroy(g.aux_done[i]);
                for (int i = 0; i < Context::N_AUX; ++i) cudaEventDestroy(g.aux_last[i]);
                for (int i = 0; i <= B2R_MAX_STAGES; ++i) cudaEventDestroy(g.stage_ev[i]);
                for (int i = 0; i < N_TICKET_SLOTS; ++i) { cudaEventDestroy(g.tk[i].copy); cudaEventDestroy(g.tk[i].compute); }
                for (auto& st : g.stage) { if (st.host) cudaFreeHost(st.host); if (st.done) cudaEventDestroy(st.done); }
                cudaFreeHost(g.pinned_flags);
                cudaFreeHost(g.sticky);
                g.ready = false;
            }
        }
        delete c;
        g_ctx[d] = nullptr;
    }
    g_last_device = -1;
    t_device = -1;
    return 0;
}

int b2r_current_device(void) {
    Context* c = current_ctx();
    return c ? c->device : -1;
}

static int check_overflow_flags(const int* flags, int n_views, struct b2r_scene* sc) {
    int need_tri = 0, need_quad = 0;
    for (int i = 0; i < n_views; ++i) { need_tri = std::max(need_tri, flags[2 * i]); need_quad = std::max(need_quad, flags[2 * i + 1]); }
    const int need_sil = flags[2 * MAX_FLAG_VIEWS + 1];
    if (!need_tri && !need_quad && !need_sil) return flags[2 * MAX_FLAG_VIEWS] ? fail_index() : 0;
    b2r_scene_grow_lists(sc, need_tri, need_quad, need_sil);
    return fail("tile list capacity overflow in an asynchronous render: affected frames show background only; "
                "capacity was raised, render again");
}

int b2r_sync(void) {
    CTX_OR_FAIL();
    CK(cudaSetDevice(g.device));
    CK(cudaStreamSynchronize(g.stream));
    CK(cudaStreamSynchronize(g.copy_stream));
    // device-resident renders since the last sync: the per-scene sticky flags say whether every tile list fitted
    int rc = 0;
    for (b2r_scene* sc : g.pending) rc |= b2r_scene_check_sticky(sc);
    g.pending.clear();
    return rc;
}
int64_t b2r_last_ticket(void) {
    Context* c = current_ctx();
    return c ? (int64_t)c->ticket : 0;
}

// completes the ticket held by a slot: waits for its copies and kernels (context lock dropped meanwhile when `lk`
// is given), checks the overflow flags.  Returns 0 / 1 like an entry point.
static int finish_ticket_slot(Context& g, int tslot, std::unique_lock<std::recursive_mutex>* lk) {
    Context::TicketSlot& T = g.tk[tslot];
    if (!T.open) return 0;
    const long long id = T.id;
    cudaEvent_t ev_copy = T.copy, ev_compute = T.compute;
    if (lk) lk->unlock();
    cudaError_t e1 = cudaEventSynchronize(ev_copy), e2 = cudaEventSynchronize(ev_compute);
    if (lk) lk->lock();
    if (!T.open || T.id != id) {   // someone else completed it meanwhile
        auto it = g.late_fail.find(id);
        if (it == g.late_fail.end()) return 0;
        const std::string msg = it->second;
        g.late_fail.erase(it);
        return fail(msg);
    }
    T.open = false;
    if (e1 != cudaSuccess || e2 != cudaSuccess)
        return fail(std::string("asynchronous render failed: ") + cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
    const int* flags = g.pinned_flags + (size_t)tslot * REGION_INTS;
    return check_overflow_flags(flags, T.views, T.scene);
}

int b2r_wait(int64_t ticket) {
    Context* cxp = current_ctx();
    if (!cxp) return fail("b2r_init was not called");
    Context& g = *cxp;
    std::unique_lock<std::recursive_mutex> lk(g.mu);
    if (ticket <= 0 || ticket > g.ticket) return fail("b2r_wait: unknown ticket");
    CK(cudaSetDevice(g.device));
    const int tslot = (int)(ticket % N_TICKET_SLOTS);
    if (g.tk[tslot].open && g.tk[tslot].id == ticket) return finish_ticket_slot(g, tslot, &lk);
    // already completed: by an earlier b2r_wait, or because its slot was needed again (outcome remembered)
    auto it = g.late_fail.find(ticket);
    if (it == g.late_fail.end()) return 0;
    const std::string msg = it->second;
    g.late_fail.erase(it);
    return fail(msg);
}
void* b2r_stream(void) {
    Context* c = current_ctx();
    return c ? (void*)c->stream : nullptr;
}
int64_t b2r_launch_count(void) {
    Context* c = current_ctx();
    return c ? (int64_t)c->launches : 0;
}
// Work counters of a -DB2R_STATS build (all zero otherwise): copies 16 values, optionally resets them.
int b2r_debug_stats(unsigned long long* out16, int reset) {
    CTX_OR_FAIL();
    CK(cudaSetDevice(g.device));
    CK(cudaStreamSynchronize(g.stream));
    CK(cudaMemcpyFromSymbol(out16, g_stats, sizeof(unsigned long long) * 16));
    if (reset) { unsigned long long z[16] = {0}; CK(cudaMemcpyToSymbol(g_stats, z, sizeof(z))); }
    return 0;
}
int b2r_host_alloc(int64_t bytes, void** host_ptr) {
    CTX_OR_FAIL();
    if (bytes <= 0 || !host_ptr) return fail("b2r_host_alloc: bad arguments");
    CK(cudaSetDevice(g.device));
    CK(cudaHostAlloc(host_ptr, (size_t)bytes, cudaHostAllocPortable));
    return 0;
}
int b2r_host_free(void* host_ptr) {
    if (!host_ptr) return 0;
    CK(cudaFreeHost(host_ptr));
    return 0;
}
// ---- multi-GPU output window: CUDA IPC export / import of the assembling rank's frame buffer ----
int b2r_window_create(int64_t bytes, void** dev_ptr, void* handle_out) {
    CTX_OR_FAIL();
    if (bytes <= 0 || !dev_ptr || !handle_out) return fail("b2r_window_create: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == B2R_WINDOW_HANDLE_BYTES, "IPC handle size");
    CK(cudaSetDevice(g.device));
    void* p = nullptr;
    CK(cudaMalloc(&p, (size_t)bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); return fail(std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e)); }
    std::memcpy(handle_out, &h, sizeof(h));
    *dev_ptr = p;
    return 0;
}
int b2r_window_push(const uint8_t* src_dev, uint8_t* dst_dev, int32_t n_views, int32_t height, int32_t width,
                    uint32_t* state_dev, void* stream) {
    CTX_OR_FAIL();
    if (!src_dev || !dst_dev || !state_dev || n_views <= 0 || height <= 0 || width <= 0)
        return fail("b2r_window_push: bad arguments");
    if (width % TILE_W != 0 || ((uintptr_t)src_dev | (uintptr_t)dst_dev) % 16 != 0)
        return fail("b2r_window_push: the width must be a multiple of 32 and both buffers 16-byte aligned");
    CK(cudaSetDevice(g.device));
    const int total = n_views * (width / TILE_W) * ((height + TILE_H - 1) / TILE_H);
    const int blocks = std::max(1, std::min((total + 3) / 4, 2 * g.sm_count));   // light looping CTAs (4 warps, a tile each): they share the SMs with the next render
    k_window_push<<<blocks, PUSH_THREADS, 0, (cudaStream_t)stream>>>(src_dev, dst_dev, n_views, height, width, state_dev);
    ++g.launches;
    CK(cudaGetLastError());
    return 0;
}
int b2r_window_open(const void* handle, void** dev_ptr) {
    CTX_OR_FAIL();
    if (!handle || !dev_ptr) return fail("b2r_window_open: bad arguments");
    CK(cudaSetDevice(g.device));
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, sizeof(h));
    CK(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}
int b2r_window_close(void* dev_ptr) {
    if (!dev_ptr) return 0;
    CK(cudaIpcCloseMemHandle(dev_ptr));
    return 0;
}
int b2r_window_destroy(void* dev_ptr) {
    if (!dev_ptr) return 0;
    CK(cudaFree(dev_ptr));
    return 0;
}

int b2r_set_stage_timing(int enabled) {
    CTX_OR_FAIL();
    g.timing = enabled != 0;
    return 0;
}
int b2r_last_stage_ms(const char** names, float* ms) {
    Context* cxp = current_ctx();
    if (!cxp) return 0;
    Context& g = *cxp;
    std::lock_guard<std::recursive_mutex> lock(g.mu);
    for (int i = 0; i < g.n_stage; ++i) {
        names[i] = g.stage_name[i];
        ms[i] = 0.f;
        cudaEventElapsedTime(&ms[i], g.stage_ev[i], g.stage_ev[i + 1]);
    }
    return g.n_stage;
}

static inline double load_real(const void* base, int dtype, size_t idx) {
    return dtype == B2R_F32 ? (double)((const float*)base)[idx] : ((const double*)base)[idx];
}

int b2r_scene_create(const b2r_model_desc* models, int32_t n_models, const b2r_texture_desc* textures,
                     int32_t n_textures, const b2r_cubemap_desc* skybox, b2r_scene** out_scene) {
    CTX_OR_FAIL();
    if (!out_scene) return fail("out_scene is NULL");
    CK(cudaSetDevice(g.device));
    b2r_scene* sc = new b2r_scene();
    sc->cx = &g;
    sc->n_models = n_models;
    size_t nv = 0, nt = 0, nn = 0, nf = 0, nm = 0;
    for (int i = 0; i < n_models; ++i) {
        const b2r_model_desc& m = models[i];
        if (!m.depth_test) sc->has_no_zwrite = true;
        if (!m.vertices || !m.faces || m.n_vertices <= 0) { delete sc; return fail("model without vertices/faces"); }
        nv += m.n_vertices; nt += m.uv ? m.n_uv : 0; nn += m.normals ? m.n_normals : 0; nf += m.n_faces;
        nm += std::max(1, m.n_materials);
        sc->model_faces.push_back(m.n_faces);
    }
    if (nf > (size_t)1 << 30) { delete sc; return fail("too many faces"); }
    std::vector<double4> pos(nv);
    std::vector<double2> uv(nt);
    std::vector<double> nrm(nn * 3);
    std::vector<FaceStatic> faces(nf);
    std::vector<MaterialDev> mats(nm);
    struct HalfEdge { int model, lo, hi, face, corner, rev; };
    std::vector<HalfEdge> half;
    half.reserve(nf * 3);
    std::vector<size_t> vbase(n_models);
    size_t v0 = 0, t0 = 0, n0 = 0, f0 = 0, m0 = 0;
    for (int mi = 0; mi < n_models; ++mi) {
        const b2r_model_desc& m = models[mi];
        vbase[mi] = v0;
        for (size_t i = 0; i < (size_t)m.n_vertices; ++i)
            pos[v0 + i] = make_double4(load_real(m.vertices, m.vertex_dtype, i * 4), load_real(m.vertices, m.vertex_dtype, i * 4 + 1),
                                       load_real(m.vertices, m.vertex_dtype, i * 4 + 2), load_real(m.vertices, m.vertex_dtype, i * 4 + 3));
        if (m.uv) for (size_t i = 0; i < (size_t)m.n_uv; ++i)
            uv[t0 + i] = make_double2(load_real(m.uv, m.uv_dtype, i * 3), load_real(m.uv, m.uv_dtype, i * 3 + 1));
        if (m.normals) for (size_t i = 0; i < (size_t)m.n_normals * 3; ++i) nrm[n0 * 3 + i] = load_real(m.normals, m.normal_dtype, i);
        const int nmat = std::max(1, m.n_materials);
        for (int s = 0; s < nmat; ++s) {
            MaterialDev& M = mats[m0 + s];
            b2r_material src;
            if (m.materials && s < m.n_materials) src = m.materials[s];
            else { std::memset(&src, 0, sizeof(src)); src.Kd[0] = src.Kd[1] = src.Kd[2] = 0.8; src.Ks[0] = src.Ks[1] = src.Ks[2] = 1; src.Ns = 64; src.map_Kd = src.map_Ks = src.norm = -1; }
            for (int k = 0; k < 3; ++k) { M.Kd[k] = src.Kd[k]; M.Ks255[k] = src.Ks[k] * 255; M.Kdf[k] = (float)M.Kd[k]; M.Ks255f[k] = (float)M.Ks255[k]; }
            M.Ns = src.Ns;
            M.map_Kd = src.map_Kd < n_textures ? src.map_Kd : -1;
            M.map_Ks = src.map_Ks < n_textures ? src.map_Ks : -1;
            M.norm = src.norm < n_textures ? src.norm : -1;
            M.ns_int = (src.Ns >= 0 && src.Ns <= 4096 && src.Ns == std::floor(src.Ns)) ? (int)src.Ns : -1;
            M.ns_log2 = -1; M.pad = 0;
            M.Pm = (m.materials && s < m.n_materials) ? src.Pm : 0.5;
            M.Pr = (m.materials && s < m.n_materials) ? src.Pr : 0.5;
            for (int k = 0; k < 3; ++k) M.Ka[k] = (m.materials && s < m.n_materials) ? src.Ka[k] : (k == 0 ? 0.3 : 0.0);
            for (int k = 0; k <= 12; ++k) if (M.ns_int == (1 << k)) M.ns_log2 = k;
        }
        const int base_flags = (m.clip ? FS_CLIP : 0) | (m.vertex_dtype == B2R_F32 ? FS_VTX_F32 : 0) |
                               (m.uv ? FS_HAS_UV : 0) | (m.uv && m.uv_dtype == B2R_F32 ? FS_UV_F32 : 0) |
                               (m.normals ? FS_HAS_NORMALS : 0) | (m.depth_test ? 0 : FS_NO_ZWRITE) |
                               (m.normals && m.normal_dtype == B2R_F32 ? FS_NRM_F32 : 0);
        auto wrap = [](int idx, int n) { if (idx < 0) idx += n; return (idx < 0 || idx >= n) ? 0 : idx; };
        for (size_t f = 0; f < (size_t)m.n_faces; ++f) {
            const int32_t* r = m.faces + f * 12;
            FaceStatic& F = faces[f0 + f];
            for (int c = 0; c < 3; ++c) {
                F.v[c] = (int)(v0 + wrap(r[c * 4 + 0], m.n_vertices));
                F.t[c] = m.uv ? (int)(t0 + wrap(r[c * 4 + 1], m.n_uv)) : 0;
                F.n[c] = m.normals ? (int)(n0 + wrap(r[c * 4 + 2], m.n_normals)) : 0;
            }
            int slot = r[3];
            if (slot < 0) slot += nmat;
            if (slot < 0 || slot >= nmat) slot = 0;
            F.material = (int)(m0 + slot);
            F.flags = base_flags;
            F.model = mi;
            for (int c = 0; c < 3; ++c) {  // Edge((vi[c], vi[c+1])), undirected equality on the RAW indices
                const int a = r[c * 4 + 0], b = r[((c + 1) % 3) * 4 + 0];
                half.push_back({mi, std::min(a, b), std::max(a, b), (int)(f0 + f), c, a > b ? 1 : 0});
            }
        }
        v0 += m.n_vertices; t0 += m.uv ? m.n_uv : 0; n0 += m.normals ? m.n_normals : 0; f0 += m.n_faces; m0 += nmat;
    }
    // static edge table: undirected edges -> incident (face, direction) in (face, corner) order
    std::sort(half.begin(), half.end(), [](const HalfEdge& x, const HalfEdge& y) {
        if (x.model != y.model) return x.model < y.model;
        if (x.lo != y.lo) return x.lo < y.lo;
        if (x.hi != y.hi) return x.hi < y.hi;
        if (x.face != y.face) return x.face < y.face;
        return x.corner < y.corner;
    });
    std::vector<int2> edge_v;
    std::vector<int> edge_ptr, edge_inc(half.size()), edge_model;
    for (size_t i = 0; i < half.size(); ++i) {
        const HalfEdge& h = half[i];
        if (i == 0 || h.model != half[i - 1].model || h.lo != half[i - 1].lo || h.hi != half[i - 1].hi) {
            const b2r_model_desc& m = models[h.model];
            auto wrap = [&](int idx) { if (idx < 0) idx += m.n_vertices; return (idx < 0 || idx >= m.n_vertices) ? 0 : idx; };
            edge_v.push_back(make_int2((int)(vbase[h.model] + wrap(h.lo)), (int)(vbase[h.model] + wrap(h.hi))));
            edge_ptr.push_back((int)i);
            edge_model.push_back(h.model);
        }
        edge_inc[i] = (h.face << 1) | h.rev;
    }
    edge_ptr.push_back((int)half.size());
    sc->n_faces = (int)nf;
    sc->n_vertices = (int)nv;
    sc->n_edges = (int)edge_v.size();

#define UP(buf, vec)                                                                                          \
    do {                                                                                                      \
        cudaError_t e_ = sc->buf.reserve((vec).size());                                                       \
        if (e_ == cudaSuccess && !(vec).empty())                                                              \
            e_ = cudaMemcpyAsync(sc->buf.p, (vec).data(), (vec).size() * sizeof((vec)[0]), cudaMemcpyHostToDevice, g.stream); \
        if (e_ != cudaSuccess) { std::string m_ = cudaGetErrorString(e_); b2r_scene_destroy(sc); return fail("upload " #buf ": " + m_); } \
        sc->static_bytes += (vec).size() * sizeof((vec)[0]);                                                  \
    } while (0)
    // per-face shading records: the three index levels resolved once
    std::vector<ShadeStatic> shade(nf);
    for (size_t f = 0; f < nf; ++f) {
        const FaceStatic& F = faces[f];
        ShadeStatic& R = shade[f];
        std::memset(&R, 0, sizeof(R));
        for (int c = 0; c < 3; ++c) {
            const double4 p = pos[F.v[c]];
            R.wp[c][0] = p.x; R.wp[c][1] = p.y; R.wp[c][2] = p.z;
            if (F.flags & FS_HAS_UV) { R.uu[c] = uv[F.t[c]].x; R.vv[c] = uv[F.t[c]].y; }
            if (F.flags & FS_HAS_NORMALS) for (int k = 0; k < 3; ++k) R.vn[c][k] = nrm[(size_t)F.n[c] * 3 + k];
        }
        R.material = F.material;
        R.flags = F.flags;
    }
    // float32 lighting records: tangent_() constants in the dtype the reference evaluates them in (core.py:205-213),
    // the flat unit normal (core.py:127-130, 186-187) in place of missing vertex normals
    std::vector<ShadeLite> lite(nf);
    for (size_t f = 0; f < nf; ++f) {
        const ShadeStatic& R = shade[f];
        ShadeLite& L = lite[f];
        std::memset(&L, 0, sizeof(L));
        for (int c = 0; c < 3; ++c) {
            L.uu[c] = R.uu[c]; L.vv[c] = R.vv[c];
            for (int k = 0; k < 3; ++k) { L.wp[c][k] = (float)R.wp[c][k]; L.vn[c][k] = (float)R.vn[c][k]; }
        }
        for (int k = 0; k < 3; ++k) {
            if (R.flags & FS_VTX_F32) {
                L.r0[k] = (float)R.wp[1][k] - (float)R.wp[0][k]; L.r1[k] = (float)R.wp[2][k] - (float)R.wp[0][k];
            } else {
                L.r0[k] = (float)(R.wp[1][k] - R.wp[0][k]); L.r1[k] = (float)(R.wp[2][k] - R.wp[0][k]);
            }
        }
        if (R.flags & FS_UV_F32) {
            L.du1 = (float)R.uu[1] - (float)R.uu[0]; L.du2 = (float)R.uu[2] - (float)R.uu[0];
            L.dv1 = (float)R.vv[1] - (float)R.vv[0]; L.dv2 = (float)R.vv[2] - (float)R.vv[0];
        } else {
            L.du1 = (float)(R.uu[1] - R.uu[0]); L.du2 = (float)(R.uu[2] - R.uu[0]);
            L.dv1 = (float)(R.vv[1] - R.vv[0]); L.dv2 = (float)(R.vv[2] - R.vv[0]);
        }
        if (!(R.flags & FS_HAS_NORMALS)) {
            double e0[3], e1[3], n[3];
            for (int k = 0; k < 3; ++k) { e0[k] = R.wp[1][k] - R.wp[0][k]; e1[k] = R.wp[2][k] - R.wp[0][k]; }
            n[0] = e0[1] * e1[2] - e0[2] * e1[1]; n[1] = e0[2] * e1[0] - e0[0] * e1[2]; n[2] = e0[0] * e1[1] - e0[1] * e1[0];
            double l = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
            if (l == 0) l = 1;
            for (int c = 0; c < 3; ++c) for (int k = 0; k < 3; ++k) L.vn[c][k] = (float)(n[k] / l);
        }
        L.material = R.material; L.flags = R.flags;
    }
    std::vector<int4> face_vf(nf);
    for (size_t f = 0; f < nf; ++f) face_vf[f] = make_int4(faces[f].v[0], faces[f].v[1], faces[f].v[2], faces[f].flags);
    UP(pos, pos); UP(uv, uv); UP(nrm, nrm); UP(faces, faces); UP(face_vf, face_vf); UP(shade, shade); UP(shade_lite, lite); UP(mats, mats);
    UP(edge_v, edge_v); UP(edge_ptr, edge_ptr); UP(edge_inc, edge_inc); UP(edge_model, edge_model);
    // textures: uint8 RGB -> RGBX so that one texel is one aligned 32-bit load
    std::vector<TextureDev> tex(std::max(1, n_textures));
    std::vector<std::vector<uchar4>> staging(n_textures);
    for (int i = 0; i < n_textures; ++i) {
        const b2r_texture_desc& t = textures[i];
        const size_t n = (size_t)t.height * t.width;
        staging[i].resize(n);
        for (size_t k = 0; k < n; ++k) staging[i][k] = make_uchar4(t.rgb[k * 3], t.rgb[k * 3 + 1], t.rgb[k * 3 + 2], 255);
        uchar4* d = nullptr;
        if (cudaMalloc(&d, std::max<size_t>(n, 1) * sizeof(uchar4)) != cudaSuccess) { b2r_scene_destroy(sc); return fail("texture alloc"); }
        sc->tex_data.push_back(d);
        cudaMemcpyAsync(d, staging[i].data(), n * sizeof(uchar4), cudaMemcpyHostToDevice, g.stream);
        tex[i].texels = d; tex[i].height = t.height; tex[i].width = t.width;
        tex[i].decode = t.decode == B2R_TEX_SNORM ? 1 : 0; tex[i].tangent = t.tangent;
        sc->static_bytes += n * sizeof(uchar4);
    }
    UP(tex, tex);
    std::vector<uchar4> sky;
    if (skybox && skybox->faces && skybox->size > 0) {
        const size_t n = (size_t)6 * skybox->size * skybox->size;
        sky.resize(n);
        for (size_t k = 0; k < n; ++k) sky[k] = make_uchar4(skybox->faces[k * 3], skybox->faces[k * 3 + 1], skybox->faces[k * 3 + 2], 255);
        sc->sky_size = skybox->size;
        UP(sky, sky);
    }
#undef UP
    if (sc->sil_state.reserve(std::max(1, sc->n_edges)) != cudaSuccess) { b2r_scene_destroy(sc); return fail("alloc sil_state"); }
    cudaMemsetAsync(sc->sil_state.p, 0, std::max(1, sc->n_edges), g.stream);
    cudaError_t e = cudaStreamSynchronize(g.stream);  // staging vectors die here
    if (e != cudaSuccess) { b2r_scene_destroy(sc); return fail(std::string("scene upload: ") + cudaGetErrorString(e)); }
    for (int i = 0; i < 2; ++i)
        if (cudaEventCreateWithFlags(&sc->rgb_copied[i], cudaEventDisableTiming) != cudaSuccess) { b2r_scene_destroy(sc); return fail("event create"); }
    *out_scene = sc;
    return 0;
}

int b2r_scene_destroy(b2r_scene* sc) {
    if (!sc) return 0;
    Context& g = *sc->cx;
    std::lock_guard<std::recursive_mutex> lock_(g.mu);
    if (g.ready) {
        cudaSetDevice(g.device);
        cudaStreamSynchronize(g.stream); cudaStreamSynchronize(g.copy_stream); cudaStreamSynchronize(g.fork_stream);
        for (int i = 0; i < Context::N_AUX; ++i) cudaStreamSynchronize(g.aux[i]);
        g.pending.erase(std::remove(g.pending.begin(), g.pending.end(), sc), g.pending.end());
        for (int i = 0; i < N_TICKET_SLOTS; ++i)   // everything is drained: open tickets of this scene complete here
            if (g.tk[i].open && g.tk[i].scene == sc) {
                if (finish_ticket_slot(g, i, nullptr)) g.late_fail[g.tk[i].id] = t_error;
            }
        if (sc->sticky_slot >= 0) { for (int k = 0; k < STICKY_INTS; ++k) g.sticky[STICKY_INTS * sc->sticky_slot + k] = 0; g.sticky_free.push_back(sc->sticky_slot); }
    }
    for (int i = 0; i < 2; ++i) if (sc->rgb_copied[i]) cudaEventDestroy(sc->rgb_copied[i]);
    sc->pos.release(); sc->uv.release(); sc->nrm.release(); sc->faces.release(); sc->face_vf.release(); sc->shade.release(); sc->shade_lite.release(); sc->mats.release(); sc->tex.release();
    for (uchar4* d : sc->tex_data) cudaFree(d);
    sc->sky.release(); sc->edge_v.release(); sc->edge_ptr.release(); sc->edge_inc.release(); sc->edge_model.release();
    sc->sil_state.release(); sc->facing.release(); sc->sil.release(); sc->counters.release(); sc->views.release();
    sc->tris.release(); sc->boxes.release(); sc->coop_list.release(); sc->vrec.release(); sc->vinside.release(); sc->quads.release(); sc->tile_counts.release(); sc->tile_offs.release(); sc->tri_list.release();
    sc->quad_list.release(); sc->pair_list.release(); sc->overflow.release(); sc->tile_order.release(); sc->winner.release(); sc->packed.release(); sc->split_st.release(); sc->split_ticket.release(); sc->stencil.release(); sc->zplane.release();
    sc->status.release(); sc->frame_f32.release(); sc->rgb[0].release(); sc->rgb[1].release();
    delete sc;
    return 0;
}

int b2r_scene_reset_silhouette(b2r_scene* sc) {
    if (!sc) return fail("scene is NULL");
    Context& g = *sc->cx;
    std::lock_guard<std::recursive_mutex> lock_(g.mu);
    CK(cudaSetDevice(g.device));
    CK(cudaMemsetAsync(sc->sil_state.p, 0, std::max(1, sc->n_edges), g.stream));
    return 0;
}

int64_t b2r_scene_device_bytes(const b2r_scene* sc) { return sc ? (int64_t)sc->static_bytes : 0; }

// read the persistent silhouette back: out_pairs (n_edges, 2) GLOBAL vertex ids as stored (a,b); returns count
int b2r_scene_get_silhouette(b2r_scene* sc, int32_t* out_pairs, int32_t* out_model, int32_t capacity) {
    if (!sc) return -1;
    Context& g = *sc->cx;
    std::lock_guard<std::recursive_mutex> lock_(g.mu);
    if (cudaSetDevice(g.device) != cudaSuccess) return -1;
    std::vector<int8_t> st(std::max(1, sc->n_edges));
    std::vector<int2> ev(std::max(1, sc->n_edges));
    std::vector<int> em(std::max(1, sc->n_edges));
    if (cudaStreamSynchronize(g.stream) != cudaSuccess) return -1;
    cudaMemcpy(st.data(), sc->sil_state.p, sc->n_edges, cudaMemcpyDeviceToHost);
    cudaMemcpy(ev.data(), sc->edge_v.p, sc->n_edges * sizeof(int2), cudaMemcpyDeviceToHost);
    cudaMemcpy(em.data(), sc->edge_model.p, sc->n_edges * sizeof(int), cudaMemcpyDeviceToHost);
    int n = 0;
    for (int e = 0; e < sc->n_edges; ++e) {
        if (!st[e]) continue;
        if (n < capacity) {
            out_pairs[n * 2] = st[e] == 1 ? ev[e].x : ev[e].y;
            out_pairs[n * 2 + 1] = st[e] == 1 ? ev[e].y : ev[e].x;
            out_model[n] = em[e];
        }
        ++n;
    }
    return n;
}

int b2r_scene_set_silhouette(b2r_scene* sc, const int32_t* pairs, int32_t n) {
    if (!sc || (n > 0 && !pairs)) return fail("b2r_scene_set_silhouette: bad arguments");
    Context& g = *sc->cx;
    std::lock_guard<std::recursive_mutex> lock_(g.mu);
    CK(cudaSetDevice(g.device));
    CK(cudaStreamSynchronize(g.stream));
    const int E = sc->n_edges;
    std::vector<int2> ev(std::max(1, E));
    std::vector<int8_t> st(std::max(1, E), 0);
    if (E > 0) CK(cudaMemcpy(ev.data(), sc->edge_v.p, (size_t)E * sizeof(int2), cudaMemcpyDeviceToHost));
    std::map<std::pair<int, int>, int> index;
    for (int e = 0; e < E; ++e) index[{ev[e].x, ev[e].y}] = e;
    for (int i = 0; i < n; ++i) {
        const int a = pairs[2 * i], b = pairs[2 * i + 1];
        auto it = index.find({a, b});
        if (it != index.end()) { st[it->second] = 1; continue; }   // stored as (lo, hi) of the canonical record
        it = index.find({b, a});
        if (it != index.end()) st[it->second] = 2;                 // stored reversed
    }
    if (E > 0) CK(cudaMemcpy(sc->sil_state.p, st.data(), (size_t)E, cudaMemcpyHostToDevice));
    return 0;
}

}  // extern "C"
static void b2r_scene_grow_lists(b2r_scene* sc, int need_tri, int need_quad, int need_sil) {
    if (!sc) return;
    if (need_sil) { sc->sil_cap = std::min(std::max(1, sc->n_edges), need_sil + need_sil / 4); sc->sil.release(); sc->quads.release(); }
    if (need_tri) { sc->tri_cap = need_tri + need_tri / 4; sc->tri_list.release(); }
    if (need_quad) { sc->quad_cap = need_quad + need_quad / 4; sc->quad_list.release(); sc->pair_list.release(); }
}
// Device-resident renders only enqueue work; whether their tile lists fitted is known once the stream has drained.
// k_scan leaves the required size in the scene's sticky pinned slot when a list overflows (in ANY call since the last
// check -- a later, fitting call does not erase it).  Caller holds the context lock and has synchronised the stream.
static int b2r_scene_check_sticky(b2r_scene* sc) {
    if (!sc || sc->sticky_slot < 0) return 0;
    sc->in_pending = false;
    volatile int* st = sc->cx->sticky + STICKY_INTS * sc->sticky_slot;
    const int need_tri = st[0], need_quad = st[1], bad_index = st[2], need_sil = st[3];
    st[0] = 0; st[1] = 0; st[2] = 0; st[3] = 0;
    if (!need_tri && !need_quad && !need_sil) return bad_index ? fail_index() : 0;
    b2r_scene_grow_lists(sc, need_tri, need_quad, need_sil);
    return fail("tile list capacity overflow in an asynchronous render: affected frames show background only; "
                "capacity was raised, render again");
}
extern "C" {

// ---- host-side evaluation of the view constants (same operation order as the reference) ----------------------------
static void host_vec4_mat4(const double v[4], const double* M, double out[4]) {
    for (int j = 0; j < 4; ++j) {
        double acc = v[0] * M[j];
        acc = std::fma(v[1], M[4 + j], acc);
        acc = std::fma(v[2], M[8 + j], acc);
        acc = std::fma(v[3], M[12 + j], acc);
        out[j] = acc;
    }
}

static void make_view(const b2r_view& v, bool with_sky, ViewDev& D) {
    std::memcpy(D.mvp, v.mvp, sizeof(D.mvp));
    std::memcpy(D.mvp_dbg, v.mvp_dbg, sizeof(D.mvp_dbg));
    std::memcpy(D.viewport, v.viewport, sizeof(D.viewport));
    std::memcpy(D.planes, v.planes, sizeof(D.planes));
    std::memcpy(D.cam_pos, v.cam_pos, sizeof(D.cam_pos));
    for (int k = 0; k < 3; ++k) D.cam_posf[k] = (float)v.cam_pos[k];
    D.pad_f = 0.f;
    D.zl_num = 2 * v.near_ * v.far_;  // core.py:226-228
    D.zl_sum = v.far_ + v.near_;
    D.zl_diff = v.far_ - v.near_;
    D.system = v.system;
    D.backface = v.backface_culling;
    std::memset(D.sky, 0, sizeof(D.sky));
    if (!with_sky) return;
    // fill_frame_from_skybox (cube_map.py:83-101): two NDC triangles at z = 1
    static const double corner[2][3][4] = {{{-1, 1, 1, 1}, {1, 1, 1, 1}, {-1, -1, 1, 1}},
                                           {{1, 1, 1, 1}, {1, -1, 1, 1}, {-1, -1, 1, 1}}};
    for (int t = 0; t < 2; ++t) {
        SkyTri& T = D.sky[t];
        long long a[3][2];
        for (int i = 0; i < 3; ++i) {
            double s[4], r[4];
            host_vec4_mat4(corner[t][i], v.viewport, s);
            a[i][0] = (long long)s[0];
            a[i][1] = (long long)s[1];
            host_vec4_mat4(corner[t][i], v.sky_inv, r);
            for (int k = 0; k < 3; ++k) T.rays[i][k] = r[k] / r[3];
        }
        T.ax = a[0][0]; T.ay = a[0][1];
        T.v0x = a[1][0] - a[0][0]; T.v0y = a[1][1] - a[0][1];
        T.v1x = a[2][0] - a[0][0]; T.v1y = a[2][1] - a[0][1];
        T.d00 = (float)(T.v0x * T.v0x + T.v0y * T.v0y);
        T.d01 = (float)(T.v0x * T.v1x + T.v0y * T.v1y);
        T.d11 = (float)(T.v1x * T.v1x + T.v1y * T.v1y);
        volatile float p0 = T.d00 * T.d11, p1 = T.d01 * T.d01;
        const float den = p0 - p1;
        T.ok = den != 0.0f;
        T.inv = T.ok ? 1.0f / den : 0.0f;
    }
}

int b2r_render(b2r_scene* sc, const b2r_frame_params* fp, const b2r_view* views, int32_t n_views, uint8_t* out_rgb,
               const b2r_debug_out* dbg, int32_t out_on_device) {
    if (!sc || !fp || !views || n_views <= 0 || !out_rgb) return fail("b2r_render: bad arguments");
    Context& g = *sc->cx;
    std::lock_guard<std::recursive_mutex> lock_(g.mu);
    if (!g.ready) return fail("the scene's device context was shut down");
    CK(cudaSetDevice(g.device));  // the current device is per host thread; callers may render from a worker thread
    const int H = fp->height, W = fp->width;
    if (H <= 0 || W <= 0 || H > 32000 || W > 32000) return fail("resolution out of range");
    const int row_begin = std::max(0, fp->row_begin), row_end = std::min(H, fp->row_end <= 0 ? H : fp->row_end);
    if (row_begin >= row_end) return fail("empty row band");
    const bool with_sky = fp->bg_mode == B2R_BG_CUBEMAP;
    if (with_sky && sc->sky_size == 0) return fail("cubemap background requested but the scene has no skybox");
    const bool want_status = dbg && dbg->face_status;
    const bool want_z = dbg && dbg->z;
    const bool want_f32 = dbg && dbg->frame_f32;

    FrameDev Fr;
    std::memset(&Fr, 0, sizeof(Fr));
    std::memcpy(Fr.light.position, fp->light.position, sizeof(double) * 3);
    std::memcpy(Fr.light.direction, fp->light.direction, sizeof(double) * 3);
    std::memcpy(Fr.light.color, fp->light.color, sizeof(double) * 3);
    std::memcpy(Fr.light.ambient, fp->light.ambient, sizeof(double) * 3);
    Fr.light.specular_strength = fp->light.specular_strength;
    Fr.light.constant = fp->light.constant; Fr.light.linear = fp->light.linear; Fr.light.quadratic = fp->light.quadratic;
    Fr.light.spot_cos_outer = fp->light.spot_cos_outer; Fr.light.spot_cos_inner = fp->light.spot_cos_inner;
    Fr.light.type = fp->light.type;
    for (int k = 0; k < 3; ++k) {
        Fr.lightf.position[k] = (float)fp->light.position[k]; Fr.lightf.direction[k] = (float)fp->light.direction[k];
        Fr.lightf.color[k] = (float)fp->light.color[k]; Fr.lightf.ambient[k] = (float)fp->light.ambient[k];
    }
    Fr.lightf.specular_strength = (float)fp->light.specular_strength;
    Fr.lightf.constant = (float)fp->light.constant; Fr.lightf.linear = (float)fp->light.linear; Fr.lightf.quadratic = (float)fp->light.quadratic;
    Fr.lightf.spot_cos_outer = (float)fp->light.spot_cos_outer;
    Fr.lightf.spot_inv_range = (float)(1.0 / (fp->light.spot_cos_inner - fp->light.spot_cos_outer));
    for (int k = 0; k < 3; ++k) Fr.background[k] = fp->background[k];
    Fr.bg_mode = fp->bg_mode;
    Fr.H = H; Fr.W = W; Fr.row_begin = row_begin; Fr.row_end = row_end;
    Fr.tiles_x = (W + TILE_W - 1) / TILE_W;
    Fr.tile_row0 = row_begin / TILE_H;
    Fr.tiles_y = (row_end - 1) / TILE_H - Fr.tile_row0 + 1;
    Fr.n_faces = sc->n_faces;
    Fr.sky_size = sc->sky_size;
    Fr.want_status = want_status;
    Fr.shading = fp->shading;
    if (Fr.shading < B2R_SHADE_GENERAL || Fr.shading > B2R_SHADE_PBR) return fail("unknown shading mode");
    Fr.err_flag = nullptr;  // set below, once the read-back region of this call is known
    // stencil counts are only needed under faces -- which, with a Model(depth_test=False) around, is no longer the
    // same as "pixels whose z-buffer was written": count everywhere then
    // B2R_DEBUG_SKIP_BG=1 (tests): hand out the stencil plane of the PRODUCTION path -- counts kept under faces only --
    // so that its shortcuts are compared with the oracle count by count on the covered pixels
    Fr.full_stencil = ((dbg && dbg->stencil && !g.debug_skip_bg) || sc->has_no_zwrite) ? 1 : 0;
    const int n_tiles = Fr.tiles_x * Fr.tiles_y;
    const int F = sc->n_faces, NV = std::max(1, sc->n_vertices);
    // Silhouette / quad records: far fewer edges are extruded at once than the mesh has (diablo 1 381 of 7 533, the 1M-
    // triangle torus ~3 000 of 1.5 M), so the records are sized by a capacity that grows on overflow like the tile lists
    if (sc->sil_cap == 0) sc->sil_cap = std::min(std::max(1, sc->n_edges), std::max(16384, sc->n_edges / 16));
    int E = sc->sil_cap;
    const size_t npx = (size_t)H * W;
    const SceneDev S = sc->dev();

    // Chunking.  Device-resident output: as many views per launch as the scratch budget allows.  Host output: small
    // chunks, so that the D2H copy of chunk i (copy stream) overlaps the kernels of chunk i+1 (compute stream).
    const bool fused = g.fused != 0 && fp->shading == B2R_SHADE_GENERAL;   // the alternative shaders live in the shading pass
    const bool want_planes = dbg && (dbg->winner || dbg->stencil);  // debug winner / stencil planes in HBM
    const size_t per_view = (size_t)F * (sizeof(TriRec) + sizeof(TriBox) + sizeof(int)) + (size_t)NV * 33 + (size_t)E * sizeof(QuadRec) + (want_planes ? npx * 6 : 0) +
                            (fused ? 0 : npx * 4) + (want_z ? npx * 8 : 0) + (want_f32 ? npx * 12 : 0);
    int VB = (int)std::max<size_t>(1, std::min<size_t>(n_views, g.scratch_budget / std::max<size_t>(per_view, 1)));
    VB = std::min(VB, 64);
    const bool host_out = out_on_device != 1;
    const bool host_async = out_on_device == 2 && !dbg;
    if (n_views > MAX_FLAG_VIEWS) return fail("too many views in one call (max 4096)");
    // Overflow read-back region.  Host-asynchronous renders take a ticket (slot t % 4; a slot still held by an
    // un-awaited ticket is completed first -- the host blocks here, i.e. at most four such calls are in flight);
    // synchronous host renders use their own region (checked before this call returns); device-resident renders write
    // a scratch region nobody reads and report through the scene's sticky flags at the next b2r_sync.
    int tslot = -1;
    long long ticket = 0;
    if (host_async) {
        ticket = ++g.ticket;
        tslot = (int)(ticket % N_TICKET_SLOTS);
        if (g.tk[tslot].open) {
            const long long old = g.tk[tslot].id;
            if (finish_ticket_slot(g, tslot, nullptr)) g.late_fail[old] = t_error;
            if (g.late_fail.size() > 64) g.late_fail.erase(g.late_fail.begin());
        }
    }
    int* const flags = g.pinned_flags + (size_t)(host_async ? tslot : ((host_out || dbg) ? N_TICKET_SLOTS : N_TICKET_SLOTS + 1)) * REGION_INTS;
    int* sticky = nullptr;
    if (!host_out && !dbg) {
        if (sc->sticky_slot < 0) {
            if (g.sticky_free.empty()) return fail("too many scenes with device-resident renders");
            sc->sticky_slot = g.sticky_free.back();
            g.sticky_free.pop_back();
        }
        sticky = g.sticky + STICKY_INTS * sc->sticky_slot;
    }
    // where the shader reports a texture lookup outside its map; the region's word is free again: its previous user
    // (ticket / synchronous call) has completed
    int* const err_flag = sticky ? sticky + 2 : flags + 2 * MAX_FLAG_VIEWS;
    int* const need_sil_flag = flags + 2 * MAX_FLAG_VIEWS + 1;
    if (!sticky) *err_flag = 0;
    *need_sil_flag = 0;
    const int rslot = host_out ? (int)(sc->host_seq++ & 1) : 0;
    Fr.err_flag = err_flag;  // pinned + unified addressing: the host pointer is valid on the device
    if (sc->tri_cap == 0) sc->tri_cap = g.init_tri_cap ? g.init_tri_cap : std::max(1 << 16, 4 * F + 8 * n_tiles);
    if (sc->quad_cap == 0) sc->quad_cap = g.init_quad_cap ? g.init_quad_cap : std::max(1 << 20, 32 * E);
    if (!g.init_tri_cap) sc->tri_cap = std::max(sc->tri_cap, 8 * n_tiles);

    g.n_stage = 0;
    if (g.timing) cudaEventRecord(g.stage_ev[0], g.stream);

    // ---- per-view constants: evaluated once, staged in pinned memory, one asynchronous upload for the whole call ----
    {
        Context::Staging& st = g.stage[g.stage_next++ % 4];
        if (!st.done) CK(cudaEventCreateWithFlags(&st.done, cudaEventDisableTiming));
        if (st.used) CK(cudaEventSynchronize(st.done));  // the upload that last used this slot has left host memory
        if (st.cap < (size_t)n_views) {
            if (st.host) cudaFreeHost(st.host);
            st.host = nullptr; st.cap = 0;
            CK(cudaMallocHost(&st.host, sizeof(ViewDev) * (size_t)std::max(n_views, 64)));
            st.cap = (size_t)std::max(n_views, 64);
        }
        for (int i = 0; i < n_views; ++i) make_view(views[i], with_sky, st.host[i]);
        CK(sc->views.reserve(n_views));
        static_assert(sizeof(ViewDev) % 4 == 0, "ViewDev is copied in 32-bit words");
        const size_t words = sizeof(ViewDev) / 4 * (size_t)n_views;
        k_copy_words<<<(unsigned)std::min<size_t>((words + 255) / 256, 64), 256, 0, g.stream>>>(
            reinterpret_cast<unsigned*>(sc->views.p), reinterpret_cast<const unsigned*>(st.host), words);
        ++g.launches;
        CK(cudaEventRecord(st.done, g.stream));
        st.used = true;
    }
    // Triangle set-up only needs the view constants: it forks to its own stream here and joins before binning, so it
    // overlaps the facing -> silhouette -> quad set-up chain (all of them small, latency-bound launches).
    const bool fork = !g.timing && !dbg && F > 0;
    if (fork) CK(cudaEventRecord(g.chunk_done, g.stream));

    const cudaMemcpyKind kind = !host_out ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    if (host_out && sc->rgb_used[rslot])
        CK(cudaStreamWaitEvent(g.stream, sc->rgb_copied[rslot], 0));  // the frames of the render before last have left rgb[rslot]
    const unsigned* bg_packed = nullptr;
    for (int attempt = 0; attempt < 4; ++attempt) {
        E = sc->sil_cap;   // may have grown after an overflow
        // ---- light-dependent, view-independent: facing flags and silhouette quads ----
        CK(sc->facing.reserve(F + 4));
        CK(sc->sil.reserve(E));
        CK(sc->counters.reserve(2 + sc->n_models));
        bg_packed = reinterpret_cast<const unsigned*>(sc->counters.p + 1 + sc->n_models);
        k_frame_consts<<<1, 32, 0, g.stream>>>(Fr, sc->counters.p, 1 + sc->n_models,
                                               reinterpret_cast<unsigned*>(sc->counters.p + 1 + sc->n_models));
        ++g.launches;
        if (F > 0) {
            k_facing<<<(F + 255) / 256, 256, 0, g.stream>>>(S, Fr.light, sc->facing.p);
            ++g.launches;
            if (sc->n_edges > 0) {
                k_silhouette<<<(sc->n_edges + 255) / 256, 256, 0, g.stream>>>(
                    S, Fr.light, sc->facing.p, fp->persist_silhouette ? sc->sil_state.p : nullptr, attempt == 0 ? 1 : 0,
                    sc->sil.p, E, sc->counters.p, sc->counters.p + 1, sc->edge_model.p);
                ++g.launches;
            }
        }
        stage_mark(g, "silhouette");

        CK(sc->tris.reserve((size_t)VB * F));
        CK(sc->boxes.reserve((size_t)VB * F));
        CK(sc->vrec.reserve((size_t)VB * NV));
        CK(sc->vinside.reserve((size_t)VB * NV + 8));
        CK(sc->quads.reserve((size_t)VB * E));
        CK(sc->tile_counts.reserve((size_t)VB * n_tiles * 2 + (size_t)VB * (3 + BIN_HUGE_CAP)));  // + pair / huge-face / queued-face counters and lists
        CK(sc->coop_list.reserve((size_t)VB * F));
        CK(sc->tile_offs.reserve((size_t)VB * (n_tiles + 1) * 2));
        CK(sc->tri_list.reserve((size_t)VB * sc->tri_cap));
        CK(sc->quad_list.reserve((size_t)VB * sc->quad_cap));
        CK(sc->pair_list.reserve((size_t)VB * sc->quad_cap));
        CK(sc->overflow.reserve((size_t)VB * 2));
        CK(sc->tile_order.reserve((size_t)VB * n_tiles + VB));   // + active-tile count per view
        if (want_planes) {
            CK(sc->winner.reserve((size_t)VB * npx));
            CK(sc->stencil.reserve((size_t)VB * npx));
        }
        if (!fused) CK(sc->packed.reserve((size_t)VB * npx));
        if (want_z) CK(sc->zplane.reserve((size_t)VB * npx));
        if (want_f32) CK(sc->frame_f32.reserve((size_t)VB * npx * 3));
        if (want_status) CK(sc->status.reserve((size_t)VB * F + 8));
        if (host_out) CK(sc->rgb[rslot].reserve((size_t)n_views * npx * 3));  // one region per view: copies never race

        for (int first = 0; first < n_views; first += VB) {
            const int nv = std::min(VB, n_views - first);
            const ViewDev* dviews = sc->views.p + first;
            const size_t count_words = (size_t)VB * n_tiles * 2 + (size_t)VB * (3 + BIN_HUGE_CAP);
            k_zero_words<<<(unsigned)std::min<size_t>((count_words + 255) / 256, (size_t)g.sm_count * 4), 256, 0, g.stream>>>(
                reinterpret_cast<unsigned*>(sc->tile_counts.p), count_words);
            ++g.launches;

            BinDev B;
            B.tri_count = sc->tile_counts.p; B.quad_count = sc->tile_counts.p + (size_t)VB * n_tiles;
            B.tri_off = sc->tile_offs.p; B.quad_off = sc->tile_offs.p + (size_t)VB * (n_tiles + 1);
            B.tri_list = sc->tri_list.p; B.quad_list = sc->quad_list.p;
            B.tri_cap = sc->tri_cap; B.quad_cap = sc->quad_cap; B.overflow = sc->overflow.p;
            B.share_cap = g.bin_share;
            B.order = sc->tile_order.p; B.n_active = sc->tile_order.p + (size_t)VB * n_tiles;
            B.pair_list = sc->pair_list.p; B.pair_count = sc->tile_counts.p + (size_t)VB * n_tiles * 2;
            B.huge_count = B.pair_count + VB; B.huge_list = B.huge_count + VB;
            int* const coop_count = B.huge_list + (size_t)VB * BIN_HUGE_CAP;
            uint8_t* status = want_status ? sc->status.p : nullptr;

            uint8_t* rgb_dev = (!host_out ? out_rgb : sc->rgb[rslot].p) + (size_t)first * npx * 3;
            TileOut T;
            T.rgb = rgb_dev; T.f32 = want_f32 ? sc->frame_f32.p : nullptr; T.bg_packed = bg_packed;
            T.packed = fused ? nullptr : sc->packed.p;
            T.winner = want_planes ? sc->winner.p : nullptr; T.stencil = want_planes ? sc->stencil.p : nullptr;
            T.z = want_z ? sc->zplane.p : nullptr; T.status = status;
            // later batches (and retries) overwrite records the previous raster launches were reading: the main stream
            // has joined all of them at the end of the previous batch, the forked set-up stream has to as well
            if (fork) CK(cudaEventRecord(g.chunk_done, g.stream));
            // The batch runs as a PIPELINE of parts of `pipe` views: set-up + binning of part p+1 (main stream, high
            // priority, small latency-bound launches) overlap the tile / shading kernels of part p (auxiliary streams),
            // which fill the machine.  Within a part, tile + shading (+ the copy to the host) may be cut further into
            // sub-chunks: sub-chunks are independent (per-view planes, read-only lists), they rotate over a few auxiliary
            // streams so that the tail of one launch overlaps the next and their frames leave over PCIe (copy stream)
            // while later sub-chunks render.
            const bool serial = g.timing || dbg;
            const int pipe = serial ? nv : std::min(nv, std::max(1, g.pipe_views));
            const int want_sub = !host_out ? g.dev_chunk : (host_async ? g.async_chunk : g.host_chunk);
            const int n_aux = !host_out ? g.aux_dev : g.aux_host;
            bool aux_used[Context::N_AUX] = {};
            int n_sub = 0;
            for (int p0 = 0; p0 < nv; p0 += pipe) {
                const int pv = std::min(pipe, nv - p0);
                if (F > 0) {
                    cudaStream_t ts = g.stream;
                    if (fork) {
                        CK(cudaStreamWaitEvent(g.fork_stream, g.chunk_done, 0));
                        ts = g.fork_stream;
                    }
                    k_vertex<<<dim3((NV + 127) / 128, pv), 128, 0, ts>>>(S, NV, dviews, sc->vrec.p, sc->vinside.p, p0);
                    ++g.launches;
                    k_tri_setup<<<dim3((F + 127) / 128, pv), 128, 0, ts>>>(S, NV, dviews, Fr, sc->vrec.p, sc->vinside.p,
                                                                           sc->tris.p, sc->boxes.p, status, coop_count,
                                                                           sc->coop_list.p, p0);
                    k_tri_count<<<dim3(std::max(1, std::min((F + 7) / 8, g.sm_count * 2)), pv), 256, 0, ts>>>(
                        S, dviews, sc->tris.p, sc->boxes.p, status, coop_count, sc->coop_list.p, p0);
                    g.launches += 2;
                    if (fork) CK(cudaEventRecord(g.tri_done, ts));
                }
                stage_mark(g, "tri_setup");
                const int quad_blocks = std::max(1, std::min((sc->n_edges + 63) / 64, g.sm_count * 4));
                k_quad_setup<<<dim3(quad_blocks, pv), 64, 0, g.stream>>>(sc->sil.p, sc->counters.p, dviews, Fr, sc->quads.p, E, p0);
                ++g.launches;
                stage_mark(g, "quad_setup");
                const int bin_blocks = g.bin_blocks ? g.bin_blocks : g.sm_count * 2;
                if (fork) {
                    // count pass in two halves: the quads (the long half) do not wait for triangle set-up, the triangles
                    // are counted on the set-up stream behind k_tri_count; both join before the scan
                    k_bin<false><<<dim3(bin_blocks, pv), 256, 0, g.stream>>>(Fr, sc->boxes.p, sc->quads.p, sc->counters.p, E, B, p0, 2);
                    k_bin<false><<<dim3(bin_blocks, pv), 256, 0, g.fork_stream>>>(Fr, sc->boxes.p, sc->quads.p, sc->counters.p, E, B, p0, 1);
                    ++g.launches;
                    CK(cudaEventRecord(g.tri_done, g.fork_stream));
                    CK(cudaStreamWaitEvent(g.stream, g.tri_done, 0));
                } else {
                    k_bin<false><<<dim3(bin_blocks, pv), 256, 0, g.stream>>>(Fr, sc->boxes.p, sc->quads.p, sc->counters.p, E, B, p0, 3);
                }
                k_scan<<<dim3(pv, 2), 1024, 0, g.stream>>>(Fr, B, flags + 2 * first, sticky, p0, sc->counters.p, E, need_sil_flag);
                k_order<<<pv, 1024, 0, g.stream>>>(Fr, B, p0);
                k_bin<true><<<dim3(bin_blocks, pv), 256, 0, g.stream>>>(Fr, sc->boxes.p, sc->quads.p, sc->counters.p, E, B, p0, 3);
                g.launches += 4;
                if (g.clip_elide && F > 0) {
                    k_clip_elide<<<dim3((n_tiles + 127) / 128, ELIDE_MAX_FACES, pv), 128, 0, g.stream>>>(
                        Fr, sc->boxes.p, sc->tris.p, S.pos, S.face_vf, dviews, B, p0);
                    ++g.launches;
                }
                stage_mark(g, "bin");
                const int sub = (serial || pv <= want_sub) ? pv : std::max(1, want_sub);
                const bool multi = !serial && (sub < pv || pipe < nv);   // tile / shading leave the main stream
                if (multi) CK(cudaEventRecord(g.setup_done, g.stream));
                for (int v0 = p0; v0 < p0 + pv; v0 += sub, ++n_sub) {
                    const int sv = std::min(sub, p0 + pv - v0);
                    const int ai = n_sub % n_aux;
                    cudaStream_t st = multi ? g.aux[ai] : g.stream;
                    if (multi) { CK(cudaStreamWaitEvent(st, g.setup_done, 0)); aux_used[ai] = true; }
                    if (fused) {
                        k_tile<true><<<dim3((unsigned)sv, (unsigned)n_tiles), RASTER_THREADS, 0, st>>>(
                            S, dviews, Fr, sc->tris.p, sc->quads.p, E, B, T, v0, sv);
                        ++g.launches;
                        stage_mark(g, "tile");
                    } else {
                        // a launch of one or two views is latency bound (it lasts as long as its heaviest tile): every tile's
                        // pair list is then cut over g.split_parts CTAs (k_tile<false, true>)
                        // (decided by the whole batch: its sub-chunks may run concurrently on several streams, so the scratch
                        // is indexed by the view's position in the batch)
                        const int parts = ((size_t)nv * n_tiles <= 8192 && E > 0) ? g.split_parts : 1;
                        if (parts > 1) {
                            CK(sc->split_st.reserve((size_t)nv * n_tiles * parts * TILE_PX));
                            CK(sc->split_ticket.reserve((size_t)nv * n_tiles));
                            const size_t slots = (size_t)sv * n_tiles;
                            k_zero_words<<<(unsigned)((slots + 255) / 256), 256, 0, st>>>(
                                reinterpret_cast<unsigned*>(sc->split_ticket.p + (size_t)v0 * n_tiles), slots);
                            ++g.launches;
                            TileOut TS = T;
                            TS.split_st = sc->split_st.p; TS.split_ticket = sc->split_ticket.p; TS.n_parts = parts;
                            k_tile<false, true><<<dim3((unsigned)sv, (unsigned)n_tiles, (unsigned)parts), RASTER_THREADS, 0, st>>>(
                                S, dviews, Fr, sc->tris.p, sc->quads.p, E, B, TS, v0, sv);
                        } else {
                            k_tile<false><<<dim3((unsigned)sv, (unsigned)n_tiles), RASTER_THREADS, 0, st>>>(
                                S, dviews, Fr, sc->tris.p, sc->quads.p, E, B, T, v0, sv);
                        }
                        ++g.launches;
                        stage_mark(g, "raster");
                        if (Fr.shading == B2R_SHADE_GENERAL && g.shade_f32 && !want_f32)
                            k_shade_packed<SHADE_F32><<<dim3((unsigned)sv, (unsigned)n_tiles), RASTER_THREADS, 0, st>>>(
                                S, dviews, Fr, sc->tris.p, B, T, v0, sv);
                        else if (Fr.shading == B2R_SHADE_GENERAL)
                            k_shade_packed<SHADE_F64><<<dim3((unsigned)sv, (unsigned)n_tiles), RASTER_THREADS, 0, st>>>(
                                S, dviews, Fr, sc->tris.p, B, T, v0, sv);
                        else
                            k_shade_packed<SHADE_ALT><<<dim3((unsigned)sv, (unsigned)n_tiles), RASTER_THREADS, 0, st>>>(
                                S, dviews, Fr, sc->tris.p, B, T, v0, sv);
                        ++g.launches;
                        stage_mark(g, "shade");
                    }
                    cudaEvent_t done = g.aux_done[n_sub % 16];
                    if (multi || host_out) CK(cudaEventRecord(done, st));
                    if (multi) CK(cudaEventRecord(g.aux_last[ai], st));
                    if (host_out) {  // these frames travel on the copy stream while the next sub-chunk renders
                        CK(cudaStreamWaitEvent(g.copy_stream, done, 0));
                        const size_t band_off = (size_t)(H - row_end) * W * 3, band_bytes = (size_t)(row_end - row_begin) * W * 3;
                        uint8_t* src = rgb_dev + (size_t)v0 * npx * 3;
                        uint8_t* dst = out_rgb + (size_t)(first + v0) * npx * 3;
                        if (band_bytes == npx * 3) {
                            CK(cudaMemcpyAsync(dst, src, (size_t)sv * npx * 3, kind, g.copy_stream));
                        } else {
                            for (int i = 0; i < sv; ++i)
                                CK(cudaMemcpyAsync(dst + (size_t)i * npx * 3 + band_off, src + (size_t)i * npx * 3 + band_off,
                                                   band_bytes, kind, g.copy_stream));
                        }
                    }
                }
            }
            for (int ai = 0; ai < Context::N_AUX; ++ai)      // the batch is complete on the main stream
                if (aux_used[ai]) CK(cudaStreamWaitEvent(g.stream, g.aux_last[ai], 0));
            if (want_status) {
                const size_t ns = (size_t)nv * F;
                k_status_resolve<<<(unsigned)((ns + 255) / 256), 256, 0, g.stream>>>(status, ns);
                ++g.launches;
            }
            CK(cudaGetLastError());
            if (dbg) {  // debug planes: same stream, so the next chunk cannot overwrite the scratch planes early
                if (dbg->z) CK(cudaMemcpyAsync(dbg->z + (size_t)first * npx, sc->zplane.p, sizeof(double) * nv * npx, kind, g.stream));
                if (dbg->frame_f32) CK(cudaMemcpyAsync(dbg->frame_f32 + (size_t)first * npx * 3, sc->frame_f32.p, sizeof(float) * nv * npx * 3, kind, g.stream));
                if (dbg->stencil) CK(cudaMemcpyAsync(dbg->stencil + (size_t)first * npx, sc->stencil.p, sizeof(short) * nv * npx, kind, g.stream));
                if (dbg->winner) CK(cudaMemcpyAsync(dbg->winner + (size_t)first * npx, sc->winner.p, sizeof(int) * nv * npx, kind, g.stream));
                if (dbg->face_status) CK(cudaMemcpyAsync(dbg->face_status + (size_t)first * F, sc->status.p, (size_t)nv * F, kind, g.stream));
                if (dbg->n_silhouette)
                    for (int i = 0; i < nv; ++i)
                        CK(cudaMemcpyAsync(dbg->n_silhouette + (size_t)(first + i) * sc->n_models, sc->counters.p + 1,
                                           sizeof(int) * sc->n_models, kind, g.stream));
            }
        }
        if (host_out) {
            CK(cudaEventRecord(sc->rgb_copied[rslot], g.copy_stream));
            sc->rgb_used[rslot] = true;
        }
        if (host_async) {  // frames are on their way: b2r_wait(ticket) blocks until they (and the overflow flags) landed
            Context::TicketSlot& T = g.tk[tslot];
            CK(cudaEventRecord(T.copy, g.copy_stream));
            CK(cudaEventRecord(T.compute, g.stream));
            T.id = ticket; T.views = n_views; T.scene = sc; T.open = true;
            return 0;
        }
        if (!host_out && !dbg) {  // asynchronous mode: capacity overflow is reported by the next b2r_sync
            if (!sc->in_pending) { g.pending.push_back(sc); sc->in_pending = true; }
            return 0;
        }
        CK(cudaStreamSynchronize(g.stream));
        CK(cudaStreamSynchronize(g.copy_stream));
        int need_tri = 0, need_quad = 0;
        for (int i = 0; i < n_views; ++i) { need_tri = std::max(need_tri, flags[2 * i]); need_quad = std::max(need_quad, flags[2 * i + 1]); }
        const int need_sil = *need_sil_flag;
        if (!need_tri && !need_quad && !need_sil) return *err_flag ? fail_index() : 0;
        b2r_scene_grow_lists(sc, need_tri, need_quad, need_sil);
        *need_sil_flag = 0;
    }
    return fail("tile list capacity overflow");
}

}  // extern "C"
