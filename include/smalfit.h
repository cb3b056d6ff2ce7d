/* smalfit.h -- C-ABI of libsmalfit (B200 / sm_100a).
 *
 * Drop-in boundary for the SMALify fitting hot path.  The reference has no native
 * interface (it is pure Python on torch + PyTorch3D); each entry point below names
 * the reference Python surface it replaces (paths relative to the SMALify checkout).
 * The Python binding a maintainer would add is in INTEGRATION.md and
 * smalify_b200/_cabi.py.
 *
 * Conventions
 *   - every function returns 0 on success or a negative SMALFIT_E* code; the message
 *     is available from smalfit_last_error().  Nothing throws across the ABI.
 *   - "dev" pointers are CUDA device pointers on the handle's device, "host"
 *     pointers are ordinary (ideally pinned) host memory.
 *   - work is enqueued on the cudaStream_t passed as `void* stream` (NULL = legacy
 *     default stream); the caller keeps every buffer alive until that work is done.
 *   - a handle is bound to one device and is not re-entrant; one handle per GPU.
 *   - frames are addressed as a contiguous range [frame0, frame0 + n_frames) of the
 *     sequence the handle was sized for.
 */
#ifndef SMALFIT_H_
#define SMALFIT_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SMALFIT_ABI_VERSION 4

#if defined(__GNUC__)
#define SMALFIT_API __attribute__((visibility("default")))
#else
#define SMALFIT_API
#endif

#define SMALFIT_OK 0
#define SMALFIT_EINVAL (-1)   /* bad argument */
#define SMALFIT_ECUDA (-2)    /* CUDA runtime error */
#define SMALFIT_ENOMEM (-3)
#define SMALFIT_ESTATE (-4)   /* call order (e.g. targets not set), or a sticky device-side fault (smalfit_status) */

/* bits of smalfit_status(): device-side faults, sticky until smalfit_destroy */
#define SMALFIT_STATUS_POOL_OVERFLOW 1   /* a frame needed more (face, tile) entries than the pool holds: that step's
                                            silhouette loss / gradient were INEXACT (entries were dropped) */
#define SMALFIT_STATUS_PEER_TIMEOUT 2    /* a peer rank did not arrive in an all-reduce: the kernel trapped */

#define SMALFIT_N_JOINTS 35
#define SMALFIT_N_POSE 34
#define SMALFIT_N_BETAS 20
#define SMALFIT_N_LOGSCALE 6
#define SMALFIT_N_KEYPOINTS 25
#define SMALFIT_N_MODEL_JOINTS 41

/* indices into the loss_terms[8] output (the `objs` dict of SMALFitter.forward,
 * smal_fitter/smal_fitter.py:138-175, plus the temporal sum and the total) */
#define SMALFIT_L_JOINT 0
#define SMALFIT_L_SIL 1
#define SMALFIT_L_BETAS 2
#define SMALFIT_L_POSE 3
#define SMALFIT_L_LIMIT 4     /* 0 unless smalfit_set_joint_limits was called: disabled in the reference (smal_fitter.py:146-151) */
#define SMALFIT_L_SPLAY 5
#define SMALFIT_L_TEMPORAL 6
#define SMALFIT_L_TOTAL 7

typedef struct smalfit_ctx* smalfit_t;

/* Model constants (HOST pointers, copied at create).  Replaces what
 * SMAL.__init__ (smal_model/smal_torch.py:24-96), Prior.__init__
 * (smal_fitter/priors/pose_prior_35.py:51-92) and the shape-prior block of
 * SMALFitter.__init__ (smal_fitter/smal_fitter.py:48-72) keep on the device.
 * Sparse tables are the ones smalify_b200/model_io.py::build_tables derives. */
typedef struct {
    int32_t n_verts, n_faces;
    const float* v_template;     /* [V*3] */
    const float* shapedirs;      /* [20 * V*3] */
    const int32_t* faces;        /* [F*3] */
    const int32_t* parents;      /* [35], root -1, parents precede children */
    const int32_t* scale_axis;   /* [35*3] log-scale index per joint axis or -1 */
    /* skinning weights: ELL by vertex and CSC by joint */
    const int32_t* skin_joint;   /* [V*8] */
    const float* skin_weight;    /* [V*8], 0-padded */
    const int32_t* skinT_ptr;    /* [36] */
    const int32_t* skinT_vert;
    const float* skinT_weight;
    /* joint regressor: rest joints (35) by joint / by vertex */
    const int32_t* jreg_ptr;     /* [36] */
    const int32_t* jreg_vert;
    const float* jreg_weight;
    const int32_t* jregT_ptr;    /* [V+1] */
    const int32_t* jregT_joint;
    const float* jregT_weight;
    /* model joints (35 regressed + 6 picked vertices) by joint / by vertex */
    const int32_t* mj_ptr;       /* [42] */
    const int32_t* mj_vert;
    const float* mj_weight;
    const int32_t* mjT_ptr;      /* [V+1] */
    const int32_t* mjT_joint;
    const float* mjT_weight;
    /* vertex -> incident (face*4 + corner) */
    const int32_t* v2f_ptr;      /* [V+1] */
    const int32_t* v2f_fc;
    const int32_t* keypoint_joint;   /* [25] model joint per keypoint (config.py:77-88) */
    /* priors */
    const float* pose_mean;      /* [105] */
    const float* pose_prec;      /* [105*105] row-major 'pic' */
    const float* pose_use;       /* [105] */
    int32_t shape_dim;           /* 26 (unity prior: betas + log scales) or 20 */
    const float* shape_mean;     /* [shape_dim] */
    const float* shape_prec;     /* [shape_dim*shape_dim] row-major */
} smalfit_model_t;

/* The five trainable tensors of SMALFitter (smal_fitter.py:58,61,83,86,89), as dev
 * pointers.  Rotations are indexed by absolute frame id. */
typedef struct {
    float* betas;             /* [n_shapes*20]  (n_shapes = 1: shared, as the reference) */
    float* log_beta_scales;   /* [n_shapes*6] */
    float* global_rotation;   /* [N*3] */
    float* joint_rotations;   /* [N*34*3] */
    float* trans;             /* [N*3] */
} smalfit_tensors_t;

/* ---- lifecycle --------------------------------------------------------- */
SMALFIT_API int smalfit_abi_version(void);

/* Replaces SMALFitter.__init__'s device setup (SMAL(...) + Renderer(...),
 * smal_fitter.py:101-102).  max_frames = N frames of the sequence, image_size = S. */
SMALFIT_API int smalfit_create(const smalfit_model_t* model, int device, int max_frames, int image_size,
                   smalfit_t* out);
/* The same with options (NULL = defaults = smalfit_create).  Zero-initialise, set struct_size = sizeof, fill in
 * what differs:
 *   frame_base, frame_capacity   this handle only ever runs the per-frame kernels on frames
 *                                [frame_base, frame_base + frame_capacity) of the max_frames-frame sequence (one rank of a
 *                                frame-sharded fit): targets and workspace are allocated for those frames only, O(N / ranks)
 *                                instead of O(N).  Parameter / gradient tensors keep their full [N] layout.
 *                                frame_capacity = 0: all frames.
 *   pool_entries_per_frame       capacity of the per-frame (face, 32x32 tile) pool; 0 = a heuristic for animals that
 *                                fill the crop like crop_to_silhouette's output (smal_fitter/utils.py:5-36).  A frame that
 *                                needs more raises SMALFIT_STATUS_POOL_OVERFLOW. */
typedef struct {
    int32_t struct_size;
    int32_t frame_base, frame_capacity;
    int32_t pool_entries_per_frame;
} smalfit_options_t;
SMALFIT_API int smalfit_create_ex(const smalfit_model_t* model, int device, int max_frames, int image_size,
                      const smalfit_options_t* options, smalfit_t* out);
SMALFIT_API void smalfit_destroy(smalfit_t h);
/* Sticky device-side faults (SMALFIT_STATUS_* bits), read without synchronising from a host-mapped word the kernels
 * write: a fault raised by step k is visible at the latest after the caller's next synchronisation.
 * smalfit_loss_grad / smalfit_fused_step also check it on entry and fail with SMALFIT_ESTATE once it is set. */
SMALFIT_API int smalfit_status(smalfit_t h, int* flags);
SMALFIT_API const char* smalfit_last_error(smalfit_t h);   /* h may be NULL: last create error */

/* ---- targets (self.sil_imgs / target_joints / target_visibility,
 *      smal_fitter.py:28-29,118-120) ------------------------------------- */
/* sil: [n*S*S] uint8 {0,1}; joints: [n*25*2] float (row,col); visibility: [n*25] uint8.
 * from_host != 0: pointers are host memory, copied H2D on `stream` (this is the copy
 * SMALFitter.forward repeats every call); 0: device pointers, copied D2D. */
SMALFIT_API int smalfit_set_targets(smalfit_t h, int frame0, int n_frames, const uint8_t* sil,
                        const float* joints, const uint8_t* visibility, int from_host, void* stream);
/* Pipelined uploads (not in the reference, which re-sends its targets inside every forward, smal_fitter.py:118-120):
 * the handle keeps a second, BACK set of target buffers.  smalfit_stage_targets fills it exactly like
 * smalfit_set_targets fills the front set, but on any stream -- typically a copy stream, while steps that read the
 * front set are running.  smalfit_swap_targets exchanges the two sets for every call enqueued afterwards (kernels
 * already enqueued keep the set they were launched with; a captured CUDA graph keeps the set it was captured with:
 * *front_index, 0 or 1, tells which one is current).  The caller orders the streams: the first step after a swap waits
 * for the staging copy (event), and a staging copy waits for the last step that read that set.  Stage every frame the
 * following steps read: a set holds what was last written into it.  smalfit_set_visibility writes the front set. */
SMALFIT_API int smalfit_stage_targets(smalfit_t h, int frame0, int n_frames, const uint8_t* sil,
                          const float* joints, const uint8_t* visibility, int from_host, void* stream);
SMALFIT_API int smalfit_swap_targets(smalfit_t h, int* front_index /* may be NULL */);
/* only the visibility rows (optimize_to_joints.py:98-110 rewrites them per stage) */
SMALFIT_API int smalfit_set_visibility(smalfit_t h, int frame0, int n_frames, const uint8_t* visibility,
                           int from_host, void* stream);
/* rotation masks global_mask[3], rotation_mask[34*3] (smal_fitter.py:92,97); host ptrs */
SMALFIT_API int smalfit_set_masks(smalfit_t h, const float* global_mask, const float* rotation_mask);
/* Optional joint-limit term (row 8f-4; the reference ships the limits in
 * smal_fitter/priors/joint_limits_prior.py but keeps the term commented out at smal_fitter.py:146-151):
 *   loss_terms[SMALFIT_L_LIMIT] = w_limit * mean over (B, 34, 3) of max(q - max, 0) + max(min - q, 0)
 * on the masked joint rotations, with its gradient.  min_limits / max_limits: HOST pointers to 34*3
 * floats (+-INFINITY = unbounded); both NULL disables the term again (the default: weights[4] is then
 * ignored exactly as the reference ignores w_limit). */
SMALFIT_API int smalfit_set_joint_limits(smalfit_t h, const float* min_limits, const float* max_limits);

/* Optional focal parameter (row 8f-4; the reference's camera is fixed: OpenGLPerspectiveCameras with the
 * default fov = 60 degrees at smal_fitter/p3d_renderer.py:22-23, i.e. x_ndc = f x_view / z_view with
 * f = 1/tan(30 deg)).  focal: DEVICE pointer to one float that replaces f for vertices and keypoints in
 * every later call (the caller keeps it alive); grad_focal: DEVICE pointer to one float that every
 * smalfit_loss_grad call overwrites with dL/dfocal over its frame range, or NULL.  focal = NULL restores the
 * reference camera.  Not part of smalfit_tensors_t: the five reference parameters stay as they are. */
SMALFIT_API int smalfit_set_focal(smalfit_t h, const float* focal, float* grad_focal);

/* frames_per_window[i] = number of frames in the window that frame i belongs to
 * (the B of every mean() in SMALFitter.forward); host pointer, N entries. Default N. */
SMALFIT_API int smalfit_set_windows(smalfit_t h, const int32_t* frames_per_window, int n_frames);

/* Extension (not in the reference): one shape (betas, log_beta_scales) per frame and one
 * window per frame, for batches of independent images.  params.betas is then [N*20]. */
SMALFIT_API int smalfit_set_per_frame_shapes(smalfit_t h, int enable);

/* ---- the hot path ------------------------------------------------------ */
/* One SMALFitter.forward(batch_range, weights, stage_id) + its backward
 * (smal_fitter.py:107-175 and the autograd pass of optimize_to_joints.py:136) for
 * frames [frame0, frame0+n_frames):
 *   weights[6] = (w_j2d, w_reproj, w_betas, w_pose, w_limit, w_splay)
 *   grads     : dL/d(each tensor), ASSIGNED (not accumulated) for the frames of the
 *               range (betas / log_beta_scales: whole tensor); a NULL member skips it
 *   loss_terms: dev float[8], see SMALFIT_L_*
 *   prior_windows: how many windows this call stands for in the shape-prior term
 *               (1 for a single forward; n_windows when a whole epoch is fused; 0 on the
 *               ranks that must not count the shared-shape prior when frames are sharded) */
SMALFIT_API int smalfit_loss_grad(smalfit_t h, const smalfit_tensors_t* params, int frame0, int n_frames,
                      const float weights[6], int prior_windows, const smalfit_tensors_t* grads,
                      float* loss_terms, void* stream);

/* One whole epoch of the reference's loop (optimize_to_joints.py:117-137) for frames [frame0, frame0 + n_frames) of
 * an n_total-frame sequence, as one kernel sequence without host involvement (CUDA-graph capturable):
 *   - loss terms + analytic gradients of every window (what smalfit_loss_grad does; window sizes from
 *     smalfit_set_windows),
 *   - get_temporal(w_temp) (smal_fitter.py:177-190) folded into the same kernels: frame i owns the pair (i, i+1) and
 *     receives the gradient of both pairs it is part of,
 *   - when peers are connected (smalfit_peer_*) and shapes are shared: every rank stores the gradient of ITS frames
 *     and its share of the shared-shape gradient / loss terms into all peers, waits for theirs, and sums in rank order
 *     -- inside the same kernel that then applies
 *   - Adam (optimize_to_joints.py:96,137; device-side step counter, reset by smalfit_adam_reset) to all frames
 *     (replicas stay bit-identical), or to the frames of the range only when nothing is exchanged.
 * grads / exp_avg / exp_avg_sq: caller-owned, full [N] layout.  loss_terms: dev float[12]: SMALFIT_L_* with
 * L_TEMPORAL filled and included in L_TOTAL, then [8..10] = the (joint, global, trans) values get_temporal returns. */
SMALFIT_API int smalfit_fused_step(smalfit_t h, const smalfit_tensors_t* params, const smalfit_tensors_t* grads,
                       const smalfit_tensors_t* exp_avg, const smalfit_tensors_t* exp_avg_sq, int frame0, int n_frames,
                       int n_total, const float weights[6], float w_temp, int prior_windows, const int32_t train[5],
                       float lr, float beta1, float beta2, float eps, float* loss_terms, void* stream);

/* SMALFitter.get_temporal(w_temp) (smal_fitter.py:177-190) over frames [0,N) and its
 * gradient ADDED into grads.{global_rotation,joint_rotations,trans}; terms = dev
 * float[3] (joint, global, trans) as the reference returns them. */
SMALFIT_API int smalfit_temporal(smalfit_t h, const smalfit_tensors_t* params, int n_frames, float w_temp,
                     const smalfit_tensors_t* grads, float* terms, void* stream);

/* torch.optim.Adam(lr, betas=(0.5,0.999), eps=1e-8).step() (optimize_to_joints.py:96,137)
 * on the tensors whose `train` flag is set; state m/v are caller-owned dev tensors of
 * the same layout (zeroed by the caller at the start of a stage).  step >= 1 uses that
 * step count; step == 0 uses a device-side counter (incremented by the call, reset by
 * smalfit_adam_reset) so that the call can be captured in a CUDA graph. */
SMALFIT_API int smalfit_adam_step(smalfit_t h, const smalfit_tensors_t* params, const smalfit_tensors_t* grads,
                      const smalfit_tensors_t* exp_avg, const smalfit_tensors_t* exp_avg_sq,
                      int n_frames, const int32_t train[5], float lr, float beta1, float beta2,
                      float eps, int step, void* stream);

SMALFIT_API int smalfit_adam_reset(smalfit_t h, void* stream);

/* ---- read-outs (Renderer.forward outputs, p3d_renderer.py:61-74) ------- */
/* soft silhouettes [n*S*S] float and projected keypoints [n*25*2] (row,col) for the
 * given frames; either output may be NULL.  No gradients. */
SMALFIT_API int smalfit_render(smalfit_t h, const smalfit_tensors_t* params, int frame0, int n_frames,
                   float* silhouettes, float* keypoints, void* stream);
/* posed vertices [n*V*3] (SMAL.__call__ verts + trans, smal_fitter.py:129) */
SMALFIT_API int smalfit_vertices(smalfit_t h, const smalfit_tensors_t* params, int frame0, int n_frames,
                     float* verts, void* stream);

/* ---- multi-GPU: one-shot all-reduce over NVLink peer memory (row 8e) ----------------------------
 * Frames are sharded over ranks (one process per GPU); per optimiser step the flat gradient (26 + 108 N floats
 * + the 8 loss terms) is summed over ranks.  These calls replace the NCCL all_reduce with one kernel per rank
 * that stores the vector into every peer's receive buffer, signals, waits for all peers and sums the slots in
 * rank order (bit-identical replicas).  Set-up, once per handle:
 *   smalfit_peer_init     allocates this rank's receive buffer for vectors of up to n_floats and returns its
 *                         64-byte CUDA IPC handle; the caller exchanges the handles (e.g. torch.distributed
 *                         all_gather) and passes all `world` of them, in rank order, to
 *   smalfit_peer_connect  which maps the peers' buffers (needs P2P access between the GPUs); synchronise the
 *                         ranks (a barrier) before the first all-reduce.
 * smalfit_peer_allreduce enqueues the kernel on `stream` (CUDA-graph capturable); every rank must call it the
 * same number of times.  A peer that does not arrive within 60 s is fatal: the kernel raises
 * SMALFIT_STATUS_PEER_TIMEOUT (smalfit_status, smalfit_peer_status) and traps, so that no rank can continue with an
 * un-reduced gradient (every later CUDA call of that process fails). */
SMALFIT_API int smalfit_peer_init(smalfit_t h, int rank, int world, int n_floats, unsigned char handle_out[64]);
SMALFIT_API int smalfit_peer_connect(smalfit_t h, const unsigned char* handles /* [world][64] */);
SMALFIT_API int smalfit_peer_allreduce(smalfit_t h, float* data, int n, void* stream);
SMALFIT_API int smalfit_peer_status(smalfit_t h, int* timed_out, void* stream);

/* ---- diagnostics ------------------------------------------------------- */
/* Per-phase device times of the most recent smalfit_loss_grad (CUDA events on the caller's
 * stream; do not enable while capturing a CUDA graph).  ms[0] shape+frame forward,
 * [1] face preparation + binning, [2] raster forward (hand-out list + tile kernel), [3] raster backward,
 * [4] frame backward, [5] shape backward + finalize, [6] whole call.
 * enable = 2 additionally makes the backward count the pairs smalfit_work_counts reports in counts[2..3] (this slows
 * the backward: do not read phase times from such a pass). */
SMALFIT_API int smalfit_set_profiling(smalfit_t h, int enable);
SMALFIT_API int smalfit_get_profile(smalfit_t h, float ms[8]);

/* Visualisation pass (row 8f-3): the reference's colour renderer (smal_fitter/p3d_renderer.py:41-59,70-72:
 * hard rasterisation, faces_per_pixel = 1, HardPhongShader, one point light at (0, 0, 3), constant vertex
 * colour, white background) of ARBITRARY world-space vertices -- generate_visualization renders the fitted
 * mesh and a copy turned by 180 degrees (smal_fitter.py:209-272).
 *   verts      DEVICE [n][V][3] float32 (e.g. from smalfit_vertices, transformed by the caller)
 *   color_rgb  HOST 3 floats in [0,1]
 *   rgb        DEVICE out [n][3][S][S] float32
 * n <= max_frames.  Uses the handle's workspace as scratch: do not overlap with a loss_grad call on another
 * stream; the backward of a preceding smalfit_loss_grad must have run before (its per-pixel buffer is reused). */
SMALFIT_API int smalfit_render_color(smalfit_t h, const float* verts, int n, const float color_rgb[3], float* rgb, void* stream);

/* counters[0] = pixels whose fragment count exceeded the K=100 cap, summed over the calls since the last read
 * counters[1] = pixels whose candidate list was longer than the 256 keys a warp selects from registers (exact, slower), same
 * counters[2] = (face, tile) entries dropped because a frame's tile pool overflowed (results INEXACT if > 0;
 *               also raises SMALFIT_STATUS_POOL_OVERFLOW), same
 * counters[3] = kernel launches since create.  counters[0..2] are reset by the call. */
SMALFIT_API int smalfit_counters(smalfit_t h, int64_t counters[4], void* stream);

/* Diagnostics (synchronises): work of the last rasterised pass over frames [frame0, frame0 + n).
 * counts[0] = (pixel, face) pairs that passed the bounding-box test (what PyTorch3D's fine rasteriser
 *             evaluates after its coarse pass; each is evaluated once in the forward and once in the backward)
 * counts[1] = (face, 32x32 tile) entries binned
 * counts[2] = of counts[0], the pairs whose pixel carries a silhouette gradient (the backward evaluates only those)
 * counts[3] = of counts[2], the pairs that are fragments selected by the K-nearest rule (contribute to the gradient).
 * counts[2..3] are accumulated by the backward while smalfit_set_profiling(h, 2) is in effect, since the last call. */
SMALFIT_API int smalfit_work_counts(smalfit_t h, int frame0, int n, int64_t counts[4], void* stream);

/* Measured FP32 ceiling of this GPU for the roofline: TFLOP/s of a register-resident FMA chain on every SM,
 * scalar FFMA (tflops[0]) and packed FFMA2 (tflops[1], fma.rn.f32x2).  Synchronises. */
SMALFIT_API int smalfit_fp32_peak(smalfit_t h, float tflops[2], void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SMALFIT_H_ */
