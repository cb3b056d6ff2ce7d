// smalfit_kernels.cuh -- kernel argument blocks shared by smalfit_kernels.cu and
// smalfit_capi.cu.  Device pointers only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "smalfit_math.cuh"

namespace smf {

constexpr int BIN_WARPS = 64;               // face segments per frame (one warp each)
constexpr int BIN_PARTS = 4;                // CTAs of a frame's cluster in frame_front
constexpr int BIN_PART_WARPS = BIN_WARPS / BIN_PARTS, BIN_PART_THREADS = BIN_PART_WARPS * 32;
constexpr int REGION_W = 8, REGION_H = 4;   // granularity of the per-row silhouette loss sums (region_l1 / region_tsum)
constexpr int TILE_W = 32, TILE_H = 32;     // rasteriser work item
constexpr int REGIONS_PER_TILE = (TILE_W / REGION_W) * (TILE_H / REGION_H);
constexpr int RT_WARPS = 8, RT_THREADS = RT_WARPS * 32;   // tile rasteriser CTA
#ifndef RT_CTAS
#define RT_CTAS 3
#endif
constexpr int RT_CTAS_PER_SM = RT_CTAS;
constexpr int RT_PITCH = TILE_W + 1;        // row pitch of a per-warp pixel plane (bank-conflict free in both directions)
constexpr int RT_PLANE = RT_PITCH * (TILE_H + 1) + 3;      // words per plane (33 x 33 corner grid of the box counts)
constexpr int RT_BLK = 16;                  // prepared faces per TMA block
constexpr int RT_BAND_MAX = 2048;           // faces of one warp's range whose positions a band item compacts (longer ranges: streamed whole)
constexpr int MAX_TILES = 1024;             // 32x32 tiles per frame (image side <= 1024)
// item-size rule of the tile rasteriser's work items (emit_items in frame_front)
constexpr int RT_FAIR = 2;                      // a hand-out item holds at most 1/RT_FAIR of a CTA's fair share of the pairs ...
constexpr int RT_MAX_ITEM = 64000;              // ... but never more pairs than this: the fragment lists of the items in flight
                                                // (n_ctas x ~0.6 x 8 B per pair) should fit the 126 MB L2 (measured at 128 frames:
                                                // 61 k pairs per item 1.05 ms / 1.06 GB of DRAM traffic, 122 k: 1.14 ms / 1.67 GB)
constexpr int RT_MIN_ITEM = 16000;              // ... and no tile is cut into bands of fewer pairs than this (an item has fixed costs;
                                                // measured best at 16 / 32 / 64 frames per GPU: 16 k / 32 k / 48 k pairs)

constexpr int MAX_LEVELS = 16;
constexpr int RT_ITEM_BINS = 24;              // size classes of the rasteriser's work items (emit_items)
constexpr int POSE_STATE_FLOATS = NJ * 42 + NLS + 3;   // R, Rw, s, t, J, G, off, theta per joint + log-scales + translation (FrameSmem's head)
constexpr int SKIN_CHUNK = 96, MAX_SKIN_CHUNKS = 192;
constexpr float P_SKIP = 2.98023224e-8f;    // 2^-25: below this 1-P rounds to 1.0f in fp32

struct ModelDev {
    int V, F, Fp, Vp;
    const float* v_template;
    const float* shapedirs;
    const ushort4* faces4;       // [Fp] (v0,v1,v2,valid) ; padding faces have valid = 0
    const int* skin_joint; const float* skin_weight;
    const int* skinT_ptr; const int* skinT_vert; const float* skinT_weight;
    // the CSC skinning weights cut into chunks of at most SKIN_CHUNK entries of one joint (frame_backward: a warp per chunk)
    int n_skin_chunks; const int* chunk_joint; const int* chunk_lo; const int* chunk_hi; const int* joint_chunk_ptr;
    const int* jreg_ptr; const int* jreg_vert; const float* jreg_weight;
    const int* jregT_ptr; const int* jregT_joint; const float* jregT_weight;
    const int* mj_ptr; const int* mj_vert; const float* mj_weight;
    const int* mjT_ptr; const int* mjT_joint; const float* mjT_weight;
    const int* v2f_ptr; const int* v2f_fc;
    const float* pose_mean; const float* pose_prec; const float* pose_use;
    int shape_dim; const float* shape_mean; const float* shape_prec;
};

// small tables in __constant__ memory
struct SkeletonConst {
    int parents[NJ];
    int scale_axis[NJ * 3];
    int joint_order[NJ];
    int level_start[MAX_LEVELS + 1];
    int n_levels;
    int child_ptr[NJ + 1];
    int child_idx[NJ];
    int kp_joint[NKP];
};

struct Params {                 // the five tensors (dev)
    const float* betas; const float* logscale; const float* glob; const float* joint; const float* trans;
};
struct Grads {
    float* betas; float* logscale; float* glob; float* joint; float* trans;
};

struct Workspace {
    int N, S, tiles_x, tiles_y;
    int n_shapes;
    int slot0;                  // workspace slot of the shared shape (= the handle's first frame)
    // per shape slot
    float* v_shaped;            // [n_shapes][V*3]
    // per frame (indexed by absolute frame id)
    float4* ndc;                // [N][Vp]  (x_ndc, y_ndc, z_view, -)
    float* gjoint;              // [N][41*3] dL/d(model joints) from the keypoint term
    float* kp_proj;             // [N][25*2]
    uint2* face_rect;           // [N][Fp]  (c0 | c1<<16, r0 | r1<<16), empty: c0 > c1
    float4* face_rec;           // [N][Fp][4] prepared faces for the backward: (x0,y0,x1,y1) (x2,y2,z0,z1)
                                //   (z2, 1/(area+eps), 1/|e01|^2, 1/|e02|^2) (1/|e12|^2, -, rect.x, rect.y)
    uint4* tile_pool;           // [N][pool_cap] binned faces (fid|v0<<16, v1|v2<<16, tile-local rect, -)
    float4* tile_rec;           // [N][pool_cap][4] the same entries as prepared faces for the tile rasteriser:
                                //   (x0,y0,x1,y1) (x2,y2,z0,z1) (z2, 1/(area+eps), 1/|e01|^2, 1/|e02|^2) (1/|e12|^2, fid, rect, -)
    unsigned* tile_off;         // [N][tiles+1] offsets into the frame's pool
    unsigned* tile_cost;        // [N][tiles] (pixel, face) pairs of the tile
    int pool_cap;
    uint2* pix;                 // [N][S*S] (float coef, u32 tkey)
    uint16_t* pix_tfid;         // [N][S*S] tie face id (capped pixels only)
    float* region_l1;           // [N][tiles*32*4] per region and pixel row: sum |alpha - T|
    float* face_grad;           // [N][Fp][8] (gx0,gy0,gx1,gy1,gx2,gy2,-,-)
    float* dvs;                 // [N][V*3]  per-frame dL/dv_shaped
    float* gw;                  // [N][V*3]  per-frame dL/d(world vertices) (frame_backward, read across its cluster)
    float* pose_state;          // [N][POSE_STATE_FLOATS] rotations, chain and skinning transforms of the frame (frame_front -> frame_backward)
    float* gJ;                  // [N][105]  per-frame dL/dJ(rest joints)
    float* gls;                 // [N][6]    per-frame dL/dlogscale
    float* frame_loss;          // [N][8]    kp, pose, splay, silhouette, joint limit, temporal (joint, global, trans)
    float* beta_partial;        // [n_shapes][n_blocks][20]
    // targets
    const uint8_t* sil;         // [N][S*S]
    const float* kp_target;     // [N][25*2]
    const uint8_t* vis;         // [N][25]
    const float* region_tsum;   // [N][tiles*32*4] per region and pixel row: sum of the target mask
    const float* inv_window;    // [N] 1 / frames_per_window
    const float* focal;         // [1] focal factor of the camera (NULL: the reference's fixed 1/tan(30 deg))
    float* gfocal;              // [1] dL/dfocal of the last loss_grad call (NULL: not wanted)
    float* gfocal_frame;        // [N][2] per-frame partials: keypoint term, silhouette term
    const float* limit_min;     // [102] joint-rotation limits (NULL: term disabled, as in the reference)
    const float* limit_max;     // [102]
    const float* gmask;         // [3]
    const float* rmask;         // [102]
    float* slot_loss;           // [N] shape-prior loss per shape slot
    unsigned* finalize_ticket;  // [1]
    float* temporal_partial;    // [blocks][3] per-block temporal loss partials
    unsigned* temporal_ticket;  // [1]
    unsigned long long* counters;   // [8] capped pixels, long lists, dropped bin entries, -, backward pairs in live pixels, backward fragments used
    unsigned* status;           // [1] host-mapped sticky fault bits (SMALFIT_STATUS_*)
    int count_pairs;            // the backward accumulates counters[4..5] (profiling)
};
constexpr unsigned STATUS_POOL_OVERFLOW = 1u, STATUS_PEER_TIMEOUT = 2u;

struct Weights {
    float j2d, sil, betas, pose, limit, splay;
    float temp;                 // w_temp of get_temporal folded into frame_backward (0: not folded, smalfit_temporal adds it)
    int n_total;                // frames of the whole sequence (temporal pairs end at n_total - 1)
    int n_terms;                // 8: loss_terms as smalfit_loss_grad documents; 12: + the three temporal values (smalfit_fused_step)
};
struct AdamState { int step; float bc1; float bc2_sqrt; int pad; };
struct AdamSegments { float* p[5]; const float* g[5]; float* m[5]; float* v[5]; int len[5]; int train[5]; };

struct TileScratch {        // tile rasteriser: per resident CTA
    uint2* list;            // [n_ctas][list_cap + Fp] fragments (depth key, 1-p) of the pixels with more than K candidates
    int list_cap;           // entries one pass may use before the tile is split into further passes
    int list_stride;        // list_cap + Fp
    unsigned* item_next;    // [1] next item to hand out
    unsigned* bin_count;    // [RT_ITEM_BINS] items per size class (class 0 = largest), filled by frame_front, cleared by the rasteriser
    unsigned* exit_ticket;  // [1] rasteriser CTAs that have finished (the last one clears the counters)
    unsigned long long* total_cost;   // [1] (pixel, face) pairs of the launch, added up tile by tile (integer atomics)
    unsigned long long* prev_total;   // [1] the previous launch's total: the item-size rule's fair share
    uint4* items;           // [RT_ITEM_BINS][bin_cap] (frame << 15 | tile << 5 | band << 2 | log2(bands), list offset, list length, -)
    unsigned bin_cap;       // = frames * tiles * 8: every item of a launch would fit one class
    unsigned short* band_idx;   // [n_ctas][RT_WARPS][RT_BAND_MAX] band items: list positions of the faces that reach the band
    int nsub;               // > 0: force this many bands for every list longer than split_len (measurements)
    int fair;               // > 0: overrides RT_FAIR (and lifts RT_MAX_ITEM)
    int split_len;          // > 0: overrides the list length above which a tile is cut into bands
    int min_item;           // > 0: overrides RT_MIN_ITEM
};

// ---- launch wrappers (defined in smalfit_kernels.cu) ----------------------
void upload_skeleton(const SkeletonConst& sk);
cudaError_t configure_kernels(const ModelDev& m);
void launch_shape_forward(const ModelDev& m, const Workspace& w, const Params& p, int frame0, int n, cudaStream_t st);
void launch_frame_front(const ModelDev& m, const Workspace& w, const TileScratch& ts, const Params& p, int frame0, int n, Weights wt,
                        float* verts_out, int do_bin /* 0: no rasteriser pass follows, 1: loss / gradient pass, 2: alpha is wanted too */,
                        int n_ctas, cudaStream_t st);
size_t raster_tile_smem_bytes();
void launch_raster_tile_forward(const ModelDev& m, const Workspace& w, const TileScratch& ts, int frame0, int n, Weights wt,
                                float* alpha_out, int n_ctas, cudaStream_t st);
constexpr int PEER_MAX = 8;                 // ranks of one NVSwitch domain
struct PeerDev {                            // one-shot all-reduce over peer-mapped memory (smalfit_peer_*)
    float* buf[PEER_MAX];                   // rank r's receive buffer: [2 parities][world][stride] floats, mapped here
    unsigned* flags[PEER_MAX];              // rank r's arrival flags: [2][world]
    unsigned* epoch;                        // [1] completed all-reduces (local)
    unsigned* ticket;                       // [1]
    unsigned* pushed;                       // [1] local CTAs that finished pushing, all epochs
    unsigned* error;                        // [1] set when a peer did not arrive within the time limit
    unsigned* status;                       // host-mapped sticky fault bits of the handle
    int rank, world, stride;
};
// The step tail (smalfit_fused_step): [exchange over peer memory] + Adam in one kernel.
struct TailArgs {
    float* p[5]; float* g[5]; float* m[5]; float* v[5];     // betas, log_beta_scales, global_rotation, joint_rotations, trans
    int train[5];
    int n_shapes, n_total;      // shapes (1: shared) and frames of the sequence
    int frame0, n_frames;       // this rank's frames
    int exchange;               // sum / gather over the connected peers (shared shapes only)
    float* terms;               // [12] loss terms of this rank (in) -> summed over ranks (out)
    float lr, b1, b2, eps;
    AdamState* state;
    unsigned* ticket;           // [1]
};
void launch_step_tail(const PeerDev& pd, const TailArgs& a, cudaStream_t st);
void launch_fp32_peak(float* out, int n_sm, int packed, int iters, cudaStream_t st);
void launch_peer_allreduce(const PeerDev& pd, float* data, int n, cudaStream_t st);
struct VisArgs;
void launch_vis_color(const ModelDev& m, const Workspace& w, const float* verts, int n, const float color[3], float focal,
                      float* rgb, cudaStream_t st);
void launch_raster_backward(const ModelDev& m, const Workspace& w, int frame0, int n, cudaStream_t st);
void launch_frame_backward(const ModelDev& m, const Workspace& w, const Params& p, const Grads& g,
                           int frame0, int n, Weights wt, cudaStream_t st);
void launch_shape_backward(const ModelDev& m, const Workspace& w, const Params& p, const Grads& g,
                           int frame0, int n, Weights wt, int prior_windows, float* loss_terms,
                           cudaStream_t st);
void launch_temporal(const Workspace& w, const Params& p, const Grads& g, int N, float w_temp,
                     float* terms, cudaStream_t st);
void launch_adam5(const AdamSegments& seg, float lr, float b1, float b2, float eps, const AdamState* s, cudaStream_t st);
void launch_adam_tick(AdamState* s, float b1, float b2, int host_step, cudaStream_t st);
void launch_adam(float* p, const float* g, float* m, float* v, int n, float lr, float b1, float b2,
                 float eps, const AdamState* s, cudaStream_t st);
void launch_region_tsum(const Workspace& w, int frame0, int n, float* region_tsum, cudaStream_t st);

}  // namespace smf
