// smalfit_capi.cu -- the C-ABI of libsmalfit (see include/smalfit.h).
// Owns the device copy of the model constants and the step workspace; every entry
// point only enqueues kernels / async copies on the caller's stream.
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "../../include/smalfit.h"
#include "smalfit_kernels.cuh"

using namespace smf;

namespace {

thread_local std::string g_create_error;

struct DevPool {                 // every allocation of a handle, freed together
    std::vector<void*> ptrs;
    cudaError_t err = cudaSuccess;
    template <typename T>
    T* alloc(size_t n, bool zero = false) {
        void* p = nullptr;
        if (err != cudaSuccess) return nullptr;
        err = cudaMalloc(&p, (n ? n : 1) * sizeof(T));
        if (err != cudaSuccess) return nullptr;
        ptrs.push_back(p);
        if (zero) err = cudaMemset(p, 0, (n ? n : 1) * sizeof(T));
        return static_cast<T*>(p);
    }
    template <typename T>
    T* upload(const T* host, size_t n) {
        T* p = alloc<T>(n);
        if (p && n) err = cudaMemcpy(p, host, n * sizeof(T), cudaMemcpyHostToDevice);
        return p;
    }
    void release() {
        for (void* p : ptrs) cudaFree(p);
        ptrs.clear();
    }
};

}  // namespace

struct smalfit_ctx {
    int device = 0;
    int N = 0, S = 0;
    int n_sm = 0;
    ModelDev m{};
    Workspace w{};
    TileScratch ts{};
    int tile_ctas = 0;
    int frame_base = 0, frame_cap = 0;   // frames this handle runs the per-frame kernels on (smalfit_options_t)
    unsigned* status_host = nullptr;     // host-mapped sticky fault word (SMALFIT_STATUS_*)
    unsigned* tail_ticket = nullptr;
    float* peak_scratch = nullptr;
    PeerDev peer{};             // one-shot all-reduce over peer memory (smalfit_peer_*); world == 0: not set up
    void* peer_local = nullptr; // this rank's allocation (receive buffer + flags + counters)
    void* peer_mapped[PEER_MAX] = {};
    size_t peer_floats = 0;
    AdamState* adam_state = nullptr;
    // mutable target buffers (Workspace holds const views)
    uint8_t* sil = nullptr; float* kp_target = nullptr; uint8_t* vis = nullptr;
    float* region_tsum = nullptr; float* inv_window = nullptr; float* gmask = nullptr; float* rmask = nullptr;
    float* limit_buf = nullptr;      // [2][102] joint-rotation limits (min, max)
    // back set of the targets (smalfit_stage_targets / smalfit_swap_targets): allocated on first use
    uint8_t* sil_back = nullptr; float* kp_back = nullptr; uint8_t* vis_back = nullptr; float* tsum_back = nullptr;
    bool back_staged = false;
    int front_index = 0;             // which of the two sets the Workspace points to (0: the one created with the handle)
    bool targets_set = false;
    DevPool pool;
    std::string error;
    long long n_launches = 0;
    bool profiling = false;
    cudaEvent_t ev[8] = {};
    bool ev_valid = false;
    void mark(int i, cudaStream_t st) { if (profiling) cudaEventRecord(ev[i], st); }
};

namespace {

int fail(smalfit_t h, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (h) h->error = buf; else g_create_error = buf;
    return code;
}

int check_cuda(smalfit_t h, cudaError_t e, const char* what) {
    if (e == cudaSuccess) return SMALFIT_OK;
    return fail(h, SMALFIT_ECUDA, "%s: %s", what, cudaGetErrorString(e));
}

int check_launch(smalfit_t h, const char* what) { return check_cuda(h, cudaGetLastError(), what); }

Params to_params(const smalfit_tensors_t* t) {
    Params p;
    p.betas = t->betas; p.logscale = t->log_beta_scales; p.glob = t->global_rotation;
    p.joint = t->joint_rotations; p.trans = t->trans;
    return p;
}
Grads to_grads(const smalfit_tensors_t* t) {
    Grads g{};
    if (t) { g.betas = t->betas; g.logscale = t->log_beta_scales; g.glob = t->global_rotation; g.joint = t->joint_rotations; g.trans = t->trans; }
    return g;
}

// per-frame kernels only run on the frames the workspace was allocated for
bool range_ok(smalfit_t h, int frame0, int n) { return n > 0 && frame0 >= h->frame_base && frame0 + n <= h->frame_base + h->frame_cap; }

int check_status(smalfit_t h, const char* who) {
    const unsigned st = h->status_host ? *(volatile unsigned*)h->status_host : 0u;
    if (st & SMALFIT_STATUS_PEER_TIMEOUT) return fail(h, SMALFIT_ESTATE, "%s: a peer rank did not arrive in an earlier all-reduce (fatal)", who);
    if (st & SMALFIT_STATUS_POOL_OVERFLOW)
        return fail(h, SMALFIT_ESTATE, "%s: an earlier step needed more (face, tile) entries than the pool holds (%d per frame) and its "
                    "silhouette loss / gradient were inexact; recreate the handle with a larger smalfit_options_t.pool_entries_per_frame",
                    who, h->w.pool_cap);
    return SMALFIT_OK;
}

}  // namespace

extern "C" {

int smalfit_abi_version(void) { return SMALFIT_ABI_VERSION; }

const char* smalfit_last_error(smalfit_t h) { return h ? h->error.c_str() : g_create_error.c_str(); }

int smalfit_create(const smalfit_model_t* md, int device, int max_frames, int image_size, smalfit_t* out) {
    return smalfit_create_ex(md, device, max_frames, image_size, nullptr, out);
}

int smalfit_create_ex(const smalfit_model_t* md, int device, int max_frames, int image_size, const smalfit_options_t* opt,
                      smalfit_t* out) {
    if (!md || !out || max_frames <= 0 || image_size <= 0 || image_size > 1024)
        return fail(nullptr, SMALFIT_EINVAL, "smalfit_create: bad arguments");
    smalfit_options_t o{};
    if (opt) {
        if (opt->struct_size < (int)sizeof(int32_t) || opt->struct_size > (int)sizeof(smalfit_options_t))
            return fail(nullptr, SMALFIT_EINVAL, "smalfit_create_ex: options.struct_size");
        memcpy(&o, opt, (size_t)opt->struct_size);
    }
    if (o.frame_capacity == 0) { o.frame_base = 0; o.frame_capacity = max_frames; }
    if (o.frame_base < 0 || o.frame_capacity < 0 || o.frame_base + o.frame_capacity > max_frames || o.pool_entries_per_frame < 0)
        return fail(nullptr, SMALFIT_EINVAL, "smalfit_create_ex: frame range / pool size out of range");
    if (md->n_verts <= 0 || md->n_verts > 65535 || md->n_faces <= 0 || md->n_faces > 65504)
        return fail(nullptr, SMALFIT_EINVAL, "smalfit_create: mesh size out of range (V=%d F=%d)", md->n_verts, md->n_faces);
    if (md->shape_dim != 26 && md->shape_dim != 20)
        return fail(nullptr, SMALFIT_EINVAL, "smalfit_create: shape_dim must be 26 or 20");
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail(nullptr, SMALFIT_ECUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return fail(nullptr, SMALFIT_ECUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major < 10)
        return fail(nullptr, SMALFIT_ECUDA, "libsmalfit is built for sm_100a; device %d is sm_%d%d", device, prop.major, prop.minor);

    smalfit_ctx* h = new smalfit_ctx();
    h->device = device; h->N = max_frames; h->S = image_size; h->n_sm = prop.multiProcessorCount;
    h->frame_base = o.frame_base; h->frame_cap = o.frame_capacity;
    const int V = md->n_verts, F = md->n_faces;
    ModelDev& m = h->m;
    m.V = V; m.F = F; m.Fp = (F + 31) / 32 * 32; m.Vp = (V + 3) / 4 * 4;
    DevPool& P = h->pool;

    // ---- skeleton tables ----
    SkeletonConst sk{};
    int depth[NJ];
    for (int j = 0; j < NJ; ++j) {
        sk.parents[j] = md->parents[j];
        if (j > 0 && (md->parents[j] < 0 || md->parents[j] >= j)) { delete h; return fail(nullptr, SMALFIT_EINVAL, "parents must precede children"); }
        depth[j] = (j == 0) ? 0 : depth[md->parents[j]] + 1;
        for (int a = 0; a < 3; ++a) sk.scale_axis[j * 3 + a] = md->scale_axis[j * 3 + a];
    }
    int maxd = 0;
    for (int j = 0; j < NJ; ++j) maxd = depth[j] > maxd ? depth[j] : maxd;
    if (maxd + 1 > MAX_LEVELS) { delete h; return fail(nullptr, SMALFIT_EINVAL, "kinematic tree too deep"); }
    sk.n_levels = maxd + 1;
    int pos = 0;
    for (int d = 0; d <= maxd; ++d) {
        sk.level_start[d] = pos;
        for (int j = 0; j < NJ; ++j) if (depth[j] == d) sk.joint_order[pos++] = j;
    }
    sk.level_start[maxd + 1] = pos;
    pos = 0;
    for (int p = 0; p < NJ; ++p) {
        sk.child_ptr[p] = pos;
        for (int j = 1; j < NJ; ++j) if (md->parents[j] == p) sk.child_idx[pos++] = j;
    }
    sk.child_ptr[NJ] = pos;
    for (int k = 0; k < NKP; ++k) sk.kp_joint[k] = md->keypoint_joint[k];
    upload_skeleton(sk);

    // ---- model constants ----
    m.v_template = P.upload(md->v_template, (size_t)V * 3);
    m.shapedirs = P.upload(md->shapedirs, (size_t)NBETA * V * 3);
    std::vector<ushort4> f4(m.Fp, make_ushort4(0, 0, 0, 0));
    for (int f = 0; f < F; ++f) {
        const int a = md->faces[f * 3], b = md->faces[f * 3 + 1], c = md->faces[f * 3 + 2];
        if (a < 0 || a >= V || b < 0 || b >= V || c < 0 || c >= V) { P.release(); delete h; return fail(nullptr, SMALFIT_EINVAL, "face index out of range"); }
        f4[f] = make_ushort4((unsigned short)a, (unsigned short)b, (unsigned short)c, 1);
    }
    m.faces4 = P.upload(f4.data(), f4.size());
    m.skin_joint = P.upload(md->skin_joint, (size_t)V * MAXINF);
    m.skin_weight = P.upload(md->skin_weight, (size_t)V * MAXINF);
    m.skinT_ptr = P.upload(md->skinT_ptr, NJ + 1);
    m.skinT_vert = P.upload(md->skinT_vert, md->skinT_ptr[NJ]);
    m.skinT_weight = P.upload(md->skinT_weight, md->skinT_ptr[NJ]);
    {   // chunks of the per-joint entry ranges (frame_backward sums a chunk per warp, then the chunks of a joint in order)
        std::vector<int> cj, clo, chi, jptr(NJ + 1, 0);
        for (int j = 0; j < NJ; ++j) {
            jptr[j] = (int)cj.size();
            for (int e = md->skinT_ptr[j]; e < md->skinT_ptr[j + 1]; e += SKIN_CHUNK) {
                cj.push_back(j); clo.push_back(e);
                chi.push_back(e + SKIN_CHUNK < md->skinT_ptr[j + 1] ? e + SKIN_CHUNK : md->skinT_ptr[j + 1]);
            }
        }
        jptr[NJ] = (int)cj.size();
        if ((int)cj.size() > MAX_SKIN_CHUNKS) { P.release(); delete h; return fail(nullptr, SMALFIT_EINVAL, "too many skinning-weight chunks (%d)", (int)cj.size()); }
        m.n_skin_chunks = (int)cj.size();
        m.chunk_joint = P.upload(cj.data(), cj.size());
        m.chunk_lo = P.upload(clo.data(), clo.size());
        m.chunk_hi = P.upload(chi.data(), chi.size());
        m.joint_chunk_ptr = P.upload(jptr.data(), jptr.size());
    }
    m.jreg_ptr = P.upload(md->jreg_ptr, NJ + 1);
    m.jreg_vert = P.upload(md->jreg_vert, md->jreg_ptr[NJ]);
    m.jreg_weight = P.upload(md->jreg_weight, md->jreg_ptr[NJ]);
    m.jregT_ptr = P.upload(md->jregT_ptr, V + 1);
    m.jregT_joint = P.upload(md->jregT_joint, md->jregT_ptr[V]);
    m.jregT_weight = P.upload(md->jregT_weight, md->jregT_ptr[V]);
    m.mj_ptr = P.upload(md->mj_ptr, NMJ + 1);
    m.mj_vert = P.upload(md->mj_vert, md->mj_ptr[NMJ]);
    m.mj_weight = P.upload(md->mj_weight, md->mj_ptr[NMJ]);
    m.mjT_ptr = P.upload(md->mjT_ptr, V + 1);
    m.mjT_joint = P.upload(md->mjT_joint, md->mjT_ptr[V]);
    m.mjT_weight = P.upload(md->mjT_weight, md->mjT_ptr[V]);
    m.v2f_ptr = P.upload(md->v2f_ptr, V + 1);
    m.v2f_fc = P.upload(md->v2f_fc, md->v2f_ptr[V]);
    m.pose_mean = P.upload(md->pose_mean, NJ * 3);
    m.pose_prec = P.upload(md->pose_prec, (size_t)NJ * 3 * NJ * 3);
    m.pose_use = P.upload(md->pose_use, NJ * 3);
    m.shape_dim = md->shape_dim;
    m.shape_mean = P.upload(md->shape_mean, md->shape_dim);
    m.shape_prec = P.upload(md->shape_prec, (size_t)md->shape_dim * md->shape_dim);

    // ---- workspace ----
    Workspace& w = h->w;
    // Per-frame buffers hold the handle's own frames only (C of them, frames [B, B + C)) but are addressed by
    // absolute frame id everywhere: their base pointers are shifted back by B frames (never dereferenced below B).
    const size_t NT = max_frames;                     // frames of the sequence: parameter-sized buffers
    const size_t N = o.frame_capacity, B = o.frame_base, SS = (size_t)image_size * image_size;
#define SHIFT(ptr, per_frame) ((ptr) ? (ptr) - B * (size_t)(per_frame) : (ptr))
    w.N = max_frames; w.S = image_size;
    w.tiles_x = (image_size + TILE_W - 1) / TILE_W; w.tiles_y = (image_size + TILE_H - 1) / TILE_H;
    const size_t tiles = (size_t)w.tiles_x * w.tiles_y;
    w.n_shapes = 1;
    w.slot0 = o.frame_base;
    const int n_blocks = (V * 3 + 255) / 256;
    w.v_shaped = SHIFT(P.alloc<float>(N * V * 3), V * 3);          // sized for per-frame shapes too (slot 0 = the shared shape)
    w.ndc = SHIFT(P.alloc<float4>(N * m.Vp), m.Vp);
    w.gjoint = SHIFT(P.alloc<float>(N * NMJ * 3, true), NMJ * 3);
    w.kp_proj = SHIFT(P.alloc<float>(N * NKP * 2, true), NKP * 2);
    w.face_rect = SHIFT(P.alloc<uint2>(N * m.Fp), m.Fp);
    w.face_rec = SHIFT(P.alloc<float4>(N * m.Fp * 4), m.Fp * 4);
    if (o.pool_entries_per_frame > 0) {
        w.pool_cap = (o.pool_entries_per_frame + 31) / 32 * 32;
    } else {   // (face, tile) entries per frame grow with the face size in pixels: ~Fp * ((bbox_px + 32) / 32)^2
        const float bbox_px = 18.f * (float)image_size / 256.f;
        const float per_face = ((bbox_px + 32.f) / 32.f) * ((bbox_px + 32.f) / 32.f);
        int mult = (int)(2.5f * per_face + 1.f);
        mult = mult < 8 ? 8 : (mult > 64 ? 64 : mult);
        w.pool_cap = mult * m.Fp;
    }
    w.tile_pool = SHIFT(P.alloc<uint4>(N * (size_t)w.pool_cap), w.pool_cap);
    w.tile_rec = SHIFT(P.alloc<float4>(N * (size_t)w.pool_cap * 4), (size_t)w.pool_cap * 4);
    {
        h->tile_ctas = h->n_sm * RT_CTAS_PER_SM;
        h->ts.list_cap = 192 * 1024;        // 1.5 MB per CTA: a 32x32 tile with ~190 candidates on every pixel in one pass
        { const char* e_cap = getenv("SMALFIT_RT_LISTCAP"); if (e_cap && atoi(e_cap) > 0) h->ts.list_cap = atoi(e_cap); }    // tests force multi-pass tiles
        h->ts.list_cap = (h->ts.list_cap + 15) / 16 * 16;            // lists start on 128-byte lines
        h->ts.list_stride = h->ts.list_cap + (m.Fp + 15) / 16 * 16 + 16;
        if ((unsigned long long)h->tile_ctas * (unsigned long long)h->ts.list_stride >= (1ull << 29)) {
            P.release(); delete h;
            return fail(nullptr, SMALFIT_EINVAL, "smalfit_create: fragment-list scratch too large for 29-bit cursors (SMALFIT_RT_LISTCAP)");
        }
        h->ts.list = P.alloc<uint2>((size_t)h->tile_ctas * h->ts.list_stride);
        h->ts.item_next = P.alloc<unsigned>(8 + RT_ITEM_BINS, true);
        h->ts.exit_ticket = h->ts.item_next + 1;
        h->ts.total_cost = reinterpret_cast<unsigned long long*>(h->ts.item_next + 2);       // (8-byte aligned: cudaMalloc base + 8)
        h->ts.prev_total = reinterpret_cast<unsigned long long*>(h->ts.item_next + 4);
        h->ts.bin_count = h->ts.item_next + 8;
        h->ts.bin_cap = (unsigned)(N * tiles * 8);
        h->ts.items = P.alloc<uint4>((size_t)RT_ITEM_BINS * h->ts.bin_cap + 8);
        h->ts.band_idx = P.alloc<unsigned short>((size_t)h->tile_ctas * RT_WARPS * RT_BAND_MAX);
        const char* e_nsub = getenv("SMALFIT_RT_NSUB");         // tuning knobs for measurements
        const char* e_split = getenv("SMALFIT_RT_SPLITLEN");
        h->ts.nsub = e_nsub ? atoi(e_nsub) : 0;
        { const char* e_fair = getenv("SMALFIT_RT_FAIR"); h->ts.fair = e_fair ? atoi(e_fair) : 0; }
        h->ts.split_len = e_split ? atoi(e_split) : 0;
        { const char* e_min = getenv("SMALFIT_RT_MINITEM"); h->ts.min_item = e_min ? atoi(e_min) : 0; }
    }
    w.tile_off = SHIFT(P.alloc<unsigned>(N * (tiles + 1), true), tiles + 1);
    w.tile_cost = SHIFT(P.alloc<unsigned>(N * tiles, true), tiles);
    w.pix = SHIFT(P.alloc<uint2>(N * SS, true), SS);
    w.pix_tfid = SHIFT(P.alloc<uint16_t>(N * SS, true), SS);
    w.region_l1 = SHIFT(P.alloc<float>(N * tiles * REGIONS_PER_TILE * REGION_H, true), tiles * REGIONS_PER_TILE * REGION_H);
    w.face_grad = SHIFT(P.alloc<float>(N * m.Fp * 8, true), m.Fp * 8);
    w.dvs = SHIFT(P.alloc<float>(N * V * 3, true), V * 3);
    w.gw = SHIFT(P.alloc<float>(N * V * 3, true), V * 3);
    w.pose_state = SHIFT(P.alloc<float>(N * POSE_STATE_FLOATS, true), POSE_STATE_FLOATS);
    w.gJ = SHIFT(P.alloc<float>(N * NJ * 3, true), NJ * 3);
    w.gls = SHIFT(P.alloc<float>(N * NLS, true), NLS);
    w.frame_loss = SHIFT(P.alloc<float>(N * 8, true), 8);
    h->limit_buf = P.alloc<float>(2 * (NJ - 1) * 3, true);
    w.gfocal_frame = SHIFT(P.alloc<float>(N * 2, true), 2);
    w.beta_partial = SHIFT(P.alloc<float>(N * n_blocks * NBETA, true), n_blocks * NBETA);
    h->sil = SHIFT(P.alloc<uint8_t>(N * SS, true), SS);
    h->kp_target = SHIFT(P.alloc<float>(N * NKP * 2, true), NKP * 2);
    h->vis = SHIFT(P.alloc<uint8_t>(N * NKP, true), NKP);
    h->region_tsum = SHIFT(P.alloc<float>(N * tiles * REGIONS_PER_TILE * REGION_H, true), tiles * REGIONS_PER_TILE * REGION_H);
    std::vector<float> ones(NT > 102 ? NT : 102, 1.0f);
    std::vector<float> invw(NT, 1.0f / (float)NT);
    h->inv_window = P.upload(invw.data(), NT);
    h->gmask = P.upload(ones.data(), 3);
    h->rmask = P.upload(ones.data(), (NJ - 1) * 3);
    w.sil = h->sil; w.kp_target = h->kp_target; w.vis = h->vis; w.region_tsum = h->region_tsum;
    w.inv_window = h->inv_window; w.gmask = h->gmask; w.rmask = h->rmask;
    h->adam_state = P.alloc<AdamState>(1, true);
    w.temporal_partial = P.alloc<float>(((size_t)NT * 108 + 255) / 256 * 3 + 3, true);
    w.temporal_ticket = P.alloc<unsigned>(1, true);
    w.slot_loss = SHIFT(P.alloc<float>(N, true), 1);
    w.finalize_ticket = P.alloc<unsigned>(1, true);
    w.counters = P.alloc<unsigned long long>(8, true);
    h->tail_ticket = P.alloc<unsigned>(1, true);
    h->peak_scratch = P.alloc<float>(4, true);
#undef SHIFT
    if (P.err == cudaSuccess) {
        P.err = cudaHostAlloc(reinterpret_cast<void**>(&h->status_host), sizeof(unsigned), cudaHostAllocMapped);
        if (P.err == cudaSuccess) {
            *h->status_host = 0u;
            P.err = cudaHostGetDevicePointer(reinterpret_cast<void**>(&w.status), h->status_host, 0);
        }
    }
    if (P.err != cudaSuccess) {
        const cudaError_t pe = P.err;
        P.release();
        if (h->status_host) cudaFreeHost(h->status_host);
        delete h;
        return fail(nullptr, pe == cudaErrorMemoryAllocation ? SMALFIT_ENOMEM : SMALFIT_ECUDA,
                    "smalfit_create: device allocation/upload failed: %s", cudaGetErrorString(pe));
    }
    e = configure_kernels(m);
    if (e != cudaSuccess) { P.release(); delete h; return fail(nullptr, SMALFIT_ECUDA, "kernel configuration: %s", cudaGetErrorString(e)); }
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { P.release(); delete h; return fail(nullptr, SMALFIT_ECUDA, "smalfit_create: %s", cudaGetErrorString(e)); }
    *out = h;
    return SMALFIT_OK;
}

void smalfit_destroy(smalfit_t h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    for (int r = 0; r < PEER_MAX; ++r) if (h->peer_mapped[r]) cudaIpcCloseMemHandle(h->peer_mapped[r]);
    if (h->peer_local) cudaFree(h->peer_local);
    h->pool.release();
    if (h->status_host) cudaFreeHost(h->status_host);
    for (int i = 0; i < 8; ++i) if (h->ev[i]) cudaEventDestroy(h->ev[i]);
    delete h;
}

int smalfit_set_per_frame_shapes(smalfit_t h, int enable) {
    if (!h) return SMALFIT_EINVAL;
    h->w.n_shapes = enable ? h->N : 1;       // slots are addressed by absolute frame id
    return SMALFIT_OK;
}

namespace {
// copies one set of targets into (sil, kp, vis) and forms the per-region row sums of the mask into tsum, all on `st`
int copy_targets(smalfit_t h, uint8_t* d_sil, float* d_kp, uint8_t* d_vis, float* d_tsum, int frame0, int n, const uint8_t* sil,
                 const float* joints, const uint8_t* visibility, int from_host, cudaStream_t st, const char* who) {
    const cudaMemcpyKind kind = from_host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    const size_t SS = (size_t)h->S * h->S;
    cudaError_t e = cudaMemcpyAsync(d_sil + frame0 * SS, sil, n * SS, kind, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_kp + (size_t)frame0 * NKP * 2, joints, (size_t)n * NKP * 2 * sizeof(float), kind, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_vis + (size_t)frame0 * NKP, visibility, (size_t)n * NKP, kind, st);
    if (e != cudaSuccess) return check_cuda(h, e, who);
    Workspace wv = h->w;
    wv.sil = d_sil;
    launch_region_tsum(wv, frame0, n, d_tsum, st);
    h->n_launches += 1;
    return check_launch(h, "region_tsum");
}
}  // namespace

int smalfit_set_targets(smalfit_t h, int frame0, int n, const uint8_t* sil, const float* joints,
                        const uint8_t* visibility, int from_host, void* stream) {
    if (!h) return SMALFIT_EINVAL;
    if (!range_ok(h, frame0, n) || !sil || !joints || !visibility) return fail(h, SMALFIT_EINVAL, "smalfit_set_targets: bad arguments");
    cudaSetDevice(h->device);
    const int rc = copy_targets(h, h->sil, h->kp_target, h->vis, h->region_tsum, frame0, n, sil, joints, visibility, from_host,
                                (cudaStream_t)stream, "smalfit_set_targets copy");
    if (rc == SMALFIT_OK) h->targets_set = true;
    return rc;
}

int smalfit_stage_targets(smalfit_t h, int frame0, int n, const uint8_t* sil, const float* joints,
                          const uint8_t* visibility, int from_host, void* stream) {
    if (!h) return SMALFIT_EINVAL;
    if (!range_ok(h, frame0, n) || !sil || !joints || !visibility) return fail(h, SMALFIT_EINVAL, "smalfit_stage_targets: bad arguments");
    cudaSetDevice(h->device);
    if (!h->sil_back) {          // second set, same shape and frame shift as the first (smalfit_create_ex)
        const size_t N = (size_t)h->frame_cap, B = (size_t)h->frame_base, SS = (size_t)h->S * h->S;
        const size_t per_t = (size_t)h->w.tiles_x * h->w.tiles_y * REGIONS_PER_TILE * REGION_H;
        uint8_t* a = h->pool.alloc<uint8_t>(N * SS, true);
        float* b = h->pool.alloc<float>(N * NKP * 2, true);
        uint8_t* c = h->pool.alloc<uint8_t>(N * NKP, true);
        float* d = h->pool.alloc<float>(N * per_t, true);
        if (h->pool.err != cudaSuccess) {
            const cudaError_t e = h->pool.err;
            h->pool.err = cudaSuccess;           // (the handle stays usable with its one set)
            return check_cuda(h, e, "smalfit_stage_targets: allocating the second set of targets");
        }
        h->sil_back = a - B * SS; h->kp_back = b - B * NKP * 2; h->vis_back = c - B * NKP; h->tsum_back = d - B * per_t;
    }
    const int rc = copy_targets(h, h->sil_back, h->kp_back, h->vis_back, h->tsum_back, frame0, n, sil, joints, visibility, from_host,
                                (cudaStream_t)stream, "smalfit_stage_targets copy");
    if (rc == SMALFIT_OK) h->back_staged = true;
    return rc;
}

int smalfit_swap_targets(smalfit_t h, int* front_index) {
    if (!h) return SMALFIT_EINVAL;
    if (!h->back_staged) return fail(h, SMALFIT_ESTATE, "smalfit_swap_targets: nothing was staged (smalfit_stage_targets)");
    std::swap(h->sil, h->sil_back); std::swap(h->kp_target, h->kp_back); std::swap(h->vis, h->vis_back);
    std::swap(h->region_tsum, h->tsum_back);
    h->w.sil = h->sil; h->w.kp_target = h->kp_target; h->w.vis = h->vis; h->w.region_tsum = h->region_tsum;
    h->back_staged = h->targets_set;     // the old front is a complete set again (if it ever was)
    h->targets_set = true;
    h->front_index ^= 1;
    if (front_index) *front_index = h->front_index;
    return SMALFIT_OK;
}

int smalfit_set_visibility(smalfit_t h, int frame0, int n, const uint8_t* visibility, int from_host, void* stream) {
    if (!h) return SMALFIT_EINVAL;
    if (!range_ok(h, frame0, n) || !visibility) return fail(h, SMALFIT_EINVAL, "smalfit_set_visibility: bad arguments");
    cudaSetDevice(h->device);
    const cudaMemcpyKind kind = from_host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    return check_cuda(h, cudaMemcpyAsync(h->vis + (size_t)frame0 * NKP, visibility, (size_t)n * NKP, kind, (cudaStream_t)stream),
                      "smalfit_set_visibility copy");
}

int smalfit_set_masks(smalfit_t h, const float* global_mask, const float* rotation_mask) {
    if (!h || !global_mask || !rotation_mask) return SMALFIT_EINVAL;
    cudaSetDevice(h->device);
    cudaError_t e = cudaMemcpy(h->gmask, global_mask, 3 * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(h->rmask, rotation_mask, (NJ - 1) * 3 * sizeof(float), cudaMemcpyHostToDevice);
    return check_cuda(h, e, "smalfit_set_masks");
}

int smalfit_set_focal(smalfit_t h, const float* focal, float* grad_focal) {
    if (!h) return SMALFIT_EINVAL;
    if (!focal && grad_focal) return fail(h, SMALFIT_EINVAL, "smalfit_set_focal: a gradient needs the parameter");
    h->w.focal = focal;
    h->w.gfocal = grad_focal;
    return SMALFIT_OK;
}

int smalfit_set_joint_limits(smalfit_t h, const float* min_limits, const float* max_limits) {
    if (!h) return SMALFIT_EINVAL;
    if (!min_limits != !max_limits) return fail(h, SMALFIT_EINVAL, "smalfit_set_joint_limits: give both limits or neither");
    if (!min_limits) { h->w.limit_min = h->w.limit_max = nullptr; return SMALFIT_OK; }
    const int n = (NJ - 1) * 3;
    for (int i = 0; i < n; ++i)
        if (!(min_limits[i] <= max_limits[i])) return fail(h, SMALFIT_EINVAL, "smalfit_set_joint_limits: min > max (or NaN) at %d", i);
    cudaSetDevice(h->device);
    cudaError_t e = cudaMemcpy(h->limit_buf, min_limits, n * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(h->limit_buf + n, max_limits, n * sizeof(float), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return check_cuda(h, e, "smalfit_set_joint_limits");
    h->w.limit_min = h->limit_buf; h->w.limit_max = h->limit_buf + n;
    return SMALFIT_OK;
}

int smalfit_set_windows(smalfit_t h, const int32_t* fpw, int n) {
    if (!h || !fpw || n <= 0 || n > h->N) return fail(h, SMALFIT_EINVAL, "smalfit_set_windows: bad arguments");
    std::vector<float> inv(n);
    for (int i = 0; i < n; ++i) {
        if (fpw[i] <= 0) return fail(h, SMALFIT_EINVAL, "smalfit_set_windows: window size must be positive");
        inv[i] = 1.0f / (float)fpw[i];
    }
    cudaSetDevice(h->device);
    return check_cuda(h, cudaMemcpy(h->inv_window, inv.data(), n * sizeof(float), cudaMemcpyHostToDevice), "smalfit_set_windows");
}

static int run_forward(smalfit_t h, const Params& p, int frame0, int n, Weights wt, bool raster, float* alpha_out,
                       float* verts_out, cudaStream_t st) {
    h->mark(0, st);
    launch_shape_forward(h->m, h->w, p, frame0, n, st);
    // frame_forward + binning of a frame in one launch over a 4-CTA cluster (profile phases 0 and 1 are reported together)
    launch_frame_front(h->m, h->w, h->ts, p, frame0, n, wt, verts_out, raster ? (alpha_out ? 2 : 1) : 0, h->tile_ctas, st);
    h->n_launches += 2;
    h->mark(1, st);
    if (raster) {
        h->mark(2, st);
        launch_raster_tile_forward(h->m, h->w, h->ts, frame0, n, wt, alpha_out, h->tile_ctas, st);
        h->n_launches += 1;
    } else {
        h->mark(2, st);
    }
    h->mark(3, st);
    return check_launch(h, "forward kernels");
}

static int run_loss_grad(smalfit_t h, const Params& p, const Grads& g, int frame0, int n, const Weights& wt, int prior_windows,
                         float* loss_terms, cudaStream_t st) {
    const bool raster = wt.sil > 0.f;
    int rc = run_forward(h, p, frame0, n, wt, raster, nullptr, nullptr, st);
    if (rc) return rc;
    if (raster) { launch_raster_backward(h->m, h->w, frame0, n, st); h->n_launches += 1; }
    h->mark(4, st);
    launch_frame_backward(h->m, h->w, p, g, frame0, n, wt, st);
    h->mark(5, st);
    launch_shape_backward(h->m, h->w, p, g, frame0, n, wt, prior_windows < 0 ? 0 : prior_windows, loss_terms, st);
    h->n_launches += 3;
    h->mark(6, st);
    if (h->profiling) h->ev_valid = true;
    return check_launch(h, "backward kernels");
}

int smalfit_loss_grad(smalfit_t h, const smalfit_tensors_t* params, int frame0, int n, const float weights[6],
                      int prior_windows, const smalfit_tensors_t* grads, float* loss_terms, void* stream) {
    if (!h) return SMALFIT_EINVAL;
    if (!params || !weights || !range_ok(h, frame0, n)) return fail(h, SMALFIT_EINVAL, "smalfit_loss_grad: bad arguments");
    if (!params->betas || !params->log_beta_scales || !params->global_rotation || !params->joint_rotations || !params->trans)
        return fail(h, SMALFIT_EINVAL, "smalfit_loss_grad: NULL parameter tensor");
    if (!h->targets_set) return fail(h, SMALFIT_ESTATE, "smalfit_loss_grad: call smalfit_set_targets first");
    if (int rc = check_status(h, "smalfit_loss_grad")) return rc;
    cudaSetDevice(h->device);
    Weights wt{weights[0], weights[1], weights[2], weights[3], weights[4], weights[5], 0.f, h->N, 8};
    return run_loss_grad(h, to_params(params), to_grads(grads), frame0, n, wt, prior_windows, loss_terms, (cudaStream_t)stream);
}

int smalfit_fused_step(smalfit_t h, const smalfit_tensors_t* params, const smalfit_tensors_t* grads, const smalfit_tensors_t* exp_avg,
                       const smalfit_tensors_t* exp_avg_sq, int frame0, int n, int n_total, const float weights[6], float w_temp,
                       int prior_windows, const int32_t train[5], float lr, float beta1, float beta2, float eps, float* loss_terms,
                       void* stream) {
    if (!h) return SMALFIT_EINVAL;
    if (!params || !grads || !exp_avg || !exp_avg_sq || !weights || !train || !loss_terms || !range_ok(h, frame0, n) ||
        n_total < frame0 + n || n_total > h->N || w_temp < 0.f)
        return fail(h, SMALFIT_EINVAL, "smalfit_fused_step: bad arguments");
    const smalfit_tensors_t* all4[4] = {params, grads, exp_avg, exp_avg_sq};
    for (const smalfit_tensors_t* t : all4)
        if (!t->betas || !t->log_beta_scales || !t->global_rotation || !t->joint_rotations || !t->trans)
            return fail(h, SMALFIT_EINVAL, "smalfit_fused_step: NULL tensor");
    if (!h->targets_set) return fail(h, SMALFIT_ESTATE, "smalfit_fused_step: call smalfit_set_targets first");
    if (int rc = check_status(h, "smalfit_fused_step")) return rc;
    const bool exchange = h->peer.epoch != nullptr && h->w.n_shapes == 1;
    if (exchange) {
        if (frame0 != h->peer.rank * n || n_total != h->peer.world * n)
            return fail(h, SMALFIT_EINVAL, "smalfit_fused_step: with peers connected rank r must own frames [r n, (r + 1) n) of world * n");
        if ((size_t)(40 + 108 * n) > h->peer_floats)
            return fail(h, SMALFIT_EINVAL, "smalfit_fused_step: peer buffers are too small for %d frames per rank", n);
    }
    cudaSetDevice(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    // the temporal term is folded into frame_backward / finalize
    Weights wt{weights[0], weights[1], weights[2], weights[3], weights[4], weights[5], w_temp, n_total, 12};
    int rc = run_loss_grad(h, to_params(params), to_grads(grads), frame0, n, wt, prior_windows, loss_terms, st);
    if (rc) return rc;
    TailArgs a{};
    const smalfit_tensors_t* src[4] = {params, grads, exp_avg, exp_avg_sq};
    float** dst[4] = {a.p, a.g, a.m, a.v};
    for (int q = 0; q < 4; ++q) {
        dst[q][0] = src[q]->betas; dst[q][1] = src[q]->log_beta_scales; dst[q][2] = src[q]->global_rotation;
        dst[q][3] = src[q]->joint_rotations; dst[q][4] = src[q]->trans;
    }
    for (int q = 0; q < 5; ++q) a.train[q] = train[q] ? 1 : 0;
    a.n_shapes = h->w.n_shapes; a.n_total = n_total; a.frame0 = frame0; a.n_frames = n; a.exchange = exchange ? 1 : 0;
    a.terms = loss_terms; a.lr = lr; a.b1 = beta1; a.b2 = beta2; a.eps = eps;
    a.state = h->adam_state; a.ticket = h->tail_ticket;
    launch_step_tail(h->peer, a, st);
    h->n_launches += 1;
    return check_launch(h, "step_tail kernel");
}

int smalfit_status(smalfit_t h, int* flags) {
    if (!h || !flags) return SMALFIT_EINVAL;
    *flags = h->status_host ? (int)*(volatile unsigned*)h->status_host : 0;
    return SMALFIT_OK;
}

int smalfit_fp32_peak(smalfit_t h, float tflops[2], void* stream) {
    if (!h || !tflops) return SMALFIT_EINVAL;
    cudaSetDevice(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    cudaEvent_t e0, e1;
    cudaError_t e = cudaEventCreate(&e0);
    if (e == cudaSuccess) e = cudaEventCreate(&e1);
    if (e != cudaSuccess) return check_cuda(h, e, "smalfit_fp32_peak");
    const int iters = 1 << 15;
    for (int packed = 0; packed < 2 && e == cudaSuccess; ++packed) {
        launch_fp32_peak(h->peak_scratch, h->n_sm, packed, 256, st);            // warm-up
        float best = 1e30f;
        for (int rep = 0; rep < 3 && e == cudaSuccess; ++rep) {
            cudaEventRecord(e0, st);
            launch_fp32_peak(h->peak_scratch, h->n_sm, packed, iters, st);
            cudaEventRecord(e1, st);
            e = cudaEventSynchronize(e1);
            float ms = 0.f;
            if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
        }
        // threads x iterations x 8 chains (4 packed pairs x 2 rounds) x 2 flop
        const double flop = (double)h->n_sm * 8 * 256 * (double)iters * 8.0 * 2.0 * (packed ? 2.0 : 1.0);
        tflops[packed] = (float)(flop / ((double)best * 1e-3) / 1e12);
    }
    h->n_launches += 8;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return check_cuda(h, e, "smalfit_fp32_peak");
}

int smalfit_temporal(smalfit_t h, const smalfit_tensors_t* params, int n_frames, float w_temp,
                     const smalfit_tensors_t* grads, float* terms, void* stream) {
    if (!h) return SMALFIT_EINVAL;
    if (!params || n_frames <= 0 || n_frames > h->N) return fail(h, SMALFIT_EINVAL, "smalfit_temporal: bad arguments");
    cudaSetDevice(h->device);
    launch_temporal(h->w, to_params(params), to_grads(grads), n_frames, w_temp, terms, (cudaStream_t)stream);
    h->n_launches += 1;
    return check_launch(h, "temporal kernel");
}

int smalfit_adam_step(smalfit_t h, const smalfit_tensors_t* params, const smalfit_tensors_t* grads,
                      const smalfit_tensors_t* exp_avg, const smalfit_tensors_t* exp_avg_sq, int n_frames,
                      const int32_t train[5], float lr, float beta1, float beta2, float eps, int step, void* stream) {
    if (!h) return SMALFIT_EINVAL;
    if (!params || !grads || !exp_avg || !exp_avg_sq || !train || step < 0 || n_frames <= 0 || n_frames > h->N)
        return fail(h, SMALFIT_EINVAL, "smalfit_adam_step: bad arguments");
    cudaSetDevice(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int ns = h->w.n_shapes;
    float* P5[5] = {params->betas, params->log_beta_scales, params->global_rotation, params->joint_rotations, params->trans};
    float* G5[5] = {grads->betas, grads->log_beta_scales, grads->global_rotation, grads->joint_rotations, grads->trans};
    float* M5[5] = {exp_avg->betas, exp_avg->log_beta_scales, exp_avg->global_rotation, exp_avg->joint_rotations, exp_avg->trans};
    float* V5[5] = {exp_avg_sq->betas, exp_avg_sq->log_beta_scales, exp_avg_sq->global_rotation, exp_avg_sq->joint_rotations, exp_avg_sq->trans};
    const int len[5] = {ns * NBETA, ns * NLS, n_frames * 3, n_frames * (NJ - 1) * 3, n_frames * 3};
    launch_adam_tick(h->adam_state, beta1, beta2, step, st);
    h->n_launches += 1;
    AdamSegments seg;
    for (int i = 0; i < 5; ++i) {
        seg.p[i] = P5[i]; seg.g[i] = G5[i]; seg.m[i] = M5[i]; seg.v[i] = V5[i]; seg.len[i] = len[i]; seg.train[i] = train[i] ? 1 : 0;
        if (train[i] && (!P5[i] || !G5[i] || !M5[i] || !V5[i])) return fail(h, SMALFIT_EINVAL, "smalfit_adam_step: NULL tensor %d", i);
    }
    launch_adam5(seg, lr, beta1, beta2, eps, h->adam_state, st);
    h->n_launches += 1;
    return check_launch(h, "adam kernel");
}

int smalfit_adam_reset(smalfit_t h, void* stream) {
    if (!h) return SMALFIT_EINVAL;
    cudaSetDevice(h->device);
    return check_cuda(h, cudaMemsetAsync(h->adam_state, 0, sizeof(AdamState), (cudaStream_t)stream), "smalfit_adam_reset");
}

int smalfit_render(smalfit_t h, const smalfit_tensors_t* params, int frame0, int n, float* silhouettes,
                   float* keypoints, void* stream) {
    if (!h) return SMALFIT_EINVAL;
    if (!params || !range_ok(h, frame0, n)) return fail(h, SMALFIT_EINVAL, "smalfit_render: bad arguments");
    cudaSetDevice(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    const Weights wt{0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, h->N, 8};
    int rc = run_forward(h, to_params(params), frame0, n, wt, silhouettes != nullptr, silhouettes, nullptr, st);
    if (rc) return rc;
    if (keypoints)
        rc = check_cuda(h, cudaMemcpyAsync(keypoints, h->w.kp_proj + (size_t)frame0 * NKP * 2, (size_t)n * NKP * 2 * sizeof(float),
                                           cudaMemcpyDeviceToDevice, st), "smalfit_render keypoints");
    return rc;
}

int smalfit_vertices(smalfit_t h, const smalfit_tensors_t* params, int frame0, int n, float* verts, void* stream) {
    if (!h) return SMALFIT_EINVAL;
    if (!params || !verts || !range_ok(h, frame0, n)) return fail(h, SMALFIT_EINVAL, "smalfit_vertices: bad arguments");
    cudaSetDevice(h->device);
    const Weights wt{0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, h->N, 8};
    return run_forward(h, to_params(params), frame0, n, wt, false, nullptr, verts, (cudaStream_t)stream);
}

int smalfit_render_color(smalfit_t h, const float* verts, int n, const float color_rgb[3], float* rgb, void* stream) {
    if (!h) return SMALFIT_EINVAL;
    if (!verts || !rgb || !color_rgb || n <= 0 || n > h->N) return fail(h, SMALFIT_EINVAL, "smalfit_render_color: bad arguments");
    cudaSetDevice(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    float focal = CAM_F;
    if (h->w.focal) {
        cudaError_t e = cudaMemcpyAsync(&focal, h->w.focal, sizeof(float), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) return check_cuda(h, e, "smalfit_render_color focal");
    }
    launch_vis_color(h->m, h->w, verts, n, color_rgb, focal, rgb, st);
    h->n_launches += 3;
    return check_launch(h, "visualisation kernels");
}

int smalfit_set_profiling(smalfit_t h, int enable) {
    if (!h) return SMALFIT_EINVAL;
    cudaSetDevice(h->device);
    if (enable && !h->ev[0]) {
        for (int i = 0; i < 8; ++i) {
            cudaError_t e = cudaEventCreate(&h->ev[i]);
            if (e != cudaSuccess) return check_cuda(h, e, "cudaEventCreate");
        }
    }
    h->profiling = enable != 0;
    h->w.count_pairs = (enable == 2) ? 1 : 0;        // the counting costs the backward ~40 %: never on in a timed pass
    h->ev_valid = false;
    return SMALFIT_OK;
}

int smalfit_get_profile(smalfit_t h, float ms[8]) {
    if (!h || !ms) return SMALFIT_EINVAL;
    if (!h->ev_valid) return fail(h, SMALFIT_ESTATE, "smalfit_get_profile: no profiled smalfit_loss_grad call yet");
    cudaSetDevice(h->device);
    cudaError_t e = cudaEventSynchronize(h->ev[6]);
    if (e != cudaSuccess) return check_cuda(h, e, "cudaEventSynchronize");
    for (int i = 0; i < 6; ++i) {
        e = cudaEventElapsedTime(&ms[i], h->ev[i], h->ev[i + 1]);
        if (e != cudaSuccess) return check_cuda(h, e, "cudaEventElapsedTime");
    }
    e = cudaEventElapsedTime(&ms[6], h->ev[0], h->ev[6]);
    ms[7] = 0.f;
    return check_cuda(h, e, "cudaEventElapsedTime");
}

// ---- one-shot all-reduce over peer memory (row 8e) ----------------------------------------------
// layout of a rank's allocation: [2][world][stride] floats | [2][world] flags | epoch | ticket | error | pushed
static size_t peer_bytes(int world, int stride) { return ((size_t)2 * world * stride) * sizeof(float) + ((size_t)2 * world + 4) * sizeof(unsigned); }

int smalfit_peer_init(smalfit_t h, int rank, int world, int n_floats, unsigned char handle_out[64]) {
    if (!h || !handle_out) return SMALFIT_EINVAL;
    if (world < 2 || world > PEER_MAX || rank < 0 || rank >= world || n_floats <= 0)
        return fail(h, SMALFIT_EINVAL, "smalfit_peer_init: need 2..%d ranks and a positive length", PEER_MAX);
    if (h->peer_local) return fail(h, SMALFIT_ESTATE, "smalfit_peer_init: already initialised");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaSetDevice(h->device);
    const int stride = (n_floats + 31) / 32 * 32;
    cudaError_t e = cudaMalloc(&h->peer_local, peer_bytes(world, stride));
    if (e == cudaSuccess) e = cudaMemset(h->peer_local, 0, peer_bytes(world, stride));
    cudaIpcMemHandle_t ih;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&ih, h->peer_local);
    if (e != cudaSuccess) { if (h->peer_local) cudaFree(h->peer_local); h->peer_local = nullptr; return check_cuda(h, e, "smalfit_peer_init"); }
    memcpy(handle_out, &ih, 64);
    h->peer.rank = rank; h->peer.world = world; h->peer.stride = stride;
    h->peer_floats = (size_t)n_floats;
    return SMALFIT_OK;
}

int smalfit_peer_connect(smalfit_t h, const unsigned char* handles) {
    if (!h || !handles) return SMALFIT_EINVAL;
    if (!h->peer_local) return fail(h, SMALFIT_ESTATE, "smalfit_peer_connect: call smalfit_peer_init first");
    cudaSetDevice(h->device);
    const int W = h->peer.world, stride = h->peer.stride;
    for (int r = 0; r < W; ++r) {
        void* base = h->peer_local;
        if (r != h->peer.rank) {
            cudaIpcMemHandle_t ih;
            memcpy(&ih, handles + (size_t)r * 64, 64);
            cudaError_t e = cudaIpcOpenMemHandle(&base, ih, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) return check_cuda(h, e, "smalfit_peer_connect: cudaIpcOpenMemHandle (NVLink / P2P access between the ranks' GPUs is required)");
            h->peer_mapped[r] = base;
        }
        h->peer.buf[r] = static_cast<float*>(base);
        h->peer.flags[r] = reinterpret_cast<unsigned*>(static_cast<float*>(base) + (size_t)2 * W * stride);
    }
    unsigned* tail = h->peer.flags[h->peer.rank] + 2 * W;
    h->peer.epoch = tail; h->peer.ticket = tail + 1; h->peer.error = tail + 2; h->peer.pushed = tail + 3;
    h->peer.status = h->w.status;
    return SMALFIT_OK;
}

int smalfit_peer_allreduce(smalfit_t h, float* data, int n, void* stream) {
    if (!h || !data) return SMALFIT_EINVAL;
    if (!h->peer_local || !h->peer.epoch) return fail(h, SMALFIT_ESTATE, "smalfit_peer_allreduce: peers are not connected");
    if (n <= 0 || (size_t)n > h->peer_floats) return fail(h, SMALFIT_EINVAL, "smalfit_peer_allreduce: length exceeds the buffer");
    cudaSetDevice(h->device);
    launch_peer_allreduce(h->peer, data, n, (cudaStream_t)stream);
    h->n_launches += 1;
    return check_launch(h, "peer_allreduce");
}

int smalfit_peer_status(smalfit_t h, int* timed_out, void* stream) {
    (void)stream;
    if (!h || !timed_out) return SMALFIT_EINVAL;
    *timed_out = (h->status_host && (*(volatile unsigned*)h->status_host & SMALFIT_STATUS_PEER_TIMEOUT)) ? 1 : 0;
    return SMALFIT_OK;
}

int smalfit_work_counts(smalfit_t h, int frame0, int n, int64_t counts[4], void* stream) {
    if (!h || !counts) return SMALFIT_EINVAL;
    if (!range_ok(h, frame0, n)) return fail(h, SMALFIT_EINVAL, "smalfit_work_counts: bad frame range");
    cudaSetDevice(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t tiles = (size_t)h->w.tiles_x * h->w.tiles_y;
    std::vector<unsigned> cost((size_t)n * tiles), off((size_t)n * (tiles + 1));
    cudaError_t e = cudaMemcpyAsync(cost.data(), h->w.tile_cost + (size_t)frame0 * tiles, cost.size() * sizeof(unsigned), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(off.data(), h->w.tile_off + (size_t)frame0 * (tiles + 1), off.size() * sizeof(unsigned), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return check_cuda(h, e, "smalfit_work_counts");
    counts[0] = counts[1] = 0;
    for (unsigned c : cost) counts[0] += c;
    for (int f = 0; f < n; ++f) counts[1] += off[(size_t)f * (tiles + 1) + tiles];
    unsigned long long bw[2] = {0, 0};
    e = cudaMemcpyAsync(bw, h->w.counters + 4, sizeof(bw), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e == cudaSuccess) e = cudaMemsetAsync(h->w.counters + 4, 0, 2 * sizeof(unsigned long long), st);
    if (e != cudaSuccess) return check_cuda(h, e, "smalfit_work_counts");
    counts[2] = (int64_t)bw[0]; counts[3] = (int64_t)bw[1];
    return SMALFIT_OK;
}

int smalfit_counters(smalfit_t h, int64_t counters[4], void* stream) {
    if (!h || !counters) return SMALFIT_EINVAL;
    cudaSetDevice(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long host[4] = {0, 0, 0, 0};
    cudaError_t e = cudaMemcpyAsync(host, h->w.counters, sizeof(host), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e == cudaSuccess) e = cudaMemsetAsync(h->w.counters, 0, 4 * sizeof(unsigned long long), st);
    if (e != cudaSuccess) return check_cuda(h, e, "smalfit_counters");
    counters[0] = (int64_t)host[0]; counters[1] = (int64_t)host[1];
    counters[2] = (int64_t)host[2]; counters[3] = h->n_launches;
    return SMALFIT_OK;
}

}  // extern "C"
