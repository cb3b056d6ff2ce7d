// smalfit_raster_tile.cuh -- forward soft rasteriser, tile formulation (included by smalfit_kernels.cu).
//
// Semantics: PyTorch3D 0.2.5 rasterize_meshes (faces_per_pixel = 100, blur 9.21e-4, no culling) +
// sigmoid_alpha_blend as configured at smal_fitter/p3d_renderer.py:26-39,66, fused with the L1
// silhouette term of smal_fitter/smal_fitter.py:172-173.  Outputs: per pixel (coef, depth threshold, tie
// face id) for raster_backward, per region row sum|alpha - T|, optionally alpha itself.
//
// Work item (emit_items: every frame's cluster in frame_front emits its own) = one 32x32-pixel tile of one frame, or one band of rows
// of a tile that holds more than the item-size rule allows (RT_FAIR / RT_MIN_ITEM / RT_MAX_ITEM); items are handed out
// largest first to persistent CTAs of 8 warps, 3 CTAs per SM.  The tile's face list (frame_front's binning, ascending
// face id) is split into 8 contiguous ranges of equal cost, one per warp, and every warp is *face-parallel*:
// it streams its prepared faces (64-byte records written by the binning) through a double-buffered
// shared-memory stage with 1-D TMA bulk copies (band items: cp.async gathers of the faces that reach the band), and
// for each face its 32 lanes sweep the face's pixel rectangle, one pixel or -- packed FP32 -- two pixels per lane.
// No pair is evaluated twice: the per-pair cost is the fragment arithmetic itself (frag_setup_forward / face_eval2,
// operation for operation the backward's face_eval_core).
//
//   P0  box counts: each warp adds its faces' rectangles into its own 33x33 corner grid (native
//       32-bit shared-memory atomics) and integrates it -> candidates per (warp, pixel).
//   P0b per pixel: c = sum over warps.  c <= K ("direct"): every fragment is selected, the pixel only needs
//       the product of (1 - p): the planes are set to 1.0f.  c > K ("listed"): the pixel gets c slots
//       (128-byte aligned) in the CTA's fragment list (global scratch); each warp's plane holds its write
//       cursor = list offset + candidates of the warps before it, so slots are in face order and the list is
//       identical from run to run.  Tiles whose lists exceed the scratch are done in several passes over
//       disjoint pixel sets (RT_SKIP).
//   P1  sweep: direct pixels multiply into the warp's plane (plain LDS/FMUL/STS: lanes of one face
//       touch distinct pixels), listed pixels store (depth key, 1 - p) at their cursor (evict_last).
//   P2a the direct pixels' products leave the planes (multiplied in warp order).
//   P2  per listed pixel, one warp: the list is staged into the warp's (now free) plane with cp.async while
//       the previous pixel is selected; keys in registers (8 per lane; lists of 257..512 entries are first narrowed
//       in the plane, longer ones in global memory), exact K-th order statistic of (depth, slot) by bisection on
//       the key bits, product over the selected set; the list's lines are then dropped from L2 without write-back.
//   P3  per pixel: alpha, |alpha - T|, coef = dL/dalpha * P / sigma; region row sums in fixed order.
// Every reduction has a fixed order: results are run-to-run deterministic.
#pragma once

namespace smf {

constexpr unsigned RT_SKIP = 0xffffffffu;       // plane value: pixel not handled in this pass
constexpr unsigned RT_LISTED = 0x80000000u;     // plane value: write cursor of a listed pixel (low 31 bits)
constexpr int RT_SELCAP = 256;                  // keys of one pixel held in registers
#ifdef RT_MIDCAP_OFF
constexpr int RT_MIDCAP = RT_SELCAP;
#else
constexpr int RT_MIDCAP = 512;                  // lists up to this length are narrowed in shared memory (a whole plane), longer ones in global memory
#endif
constexpr int RT_PIX = TILE_W * TILE_H;
constexpr unsigned char RT_CLS_DIRECT = 0, RT_CLS_LISTED = 1, RT_CLS_IDLE = 2;
constexpr int RT_MAXCHUNK = 512;                // tile lists up to 8192 faces are split by cost, longer ones evenly
constexpr unsigned RT_FACE_COST = 16u;          // per-face overhead of the sweep, in pair evaluations
// (interpolated pivots in the K-th order statistic search were measured: +2 % on the kernel -- the depths of one pixel's
//  candidates cluster on the front and back surfaces, bisection on the key bits with an exact-split exit does better)

struct RtSmem {
    unsigned plane[RT_WARPS][RT_PLANE];
    float4 stage[RT_WARPS][2][RT_BLK * 4];
    float l1[RT_PIX];
    float pl[RT_PIX];                    // product of (1 - p) per pixel, direct (plane product) or listed (selection)
    unsigned list_off[RT_PIX];           // listed pixels: list offset, then (after P2) the depth threshold
    unsigned short list_cnt[RT_PIX];     // listed pixels: slots, then (after P2) the tie face id
    unsigned short active[RT_PIX];
    unsigned long long bar[RT_WARPS][2];
    unsigned warp_tot[RT_WARPS];
    unsigned chunk[RT_MAXCHUNK];         // cost of each RT_BLK-entry chunk of the tile list, then its exclusive prefix
    unsigned char cls[RT_PIX];           // pixel class in the current pass
    unsigned cost_total;
    long long item;                      // index into ts.items of the item being processed, -1: none left
    unsigned bin_end[RT_ITEM_BINS];      // items of the size classes 0 .. b (inclusive prefix)
    unsigned n_active, total, p2_next;
    unsigned n_mid;                      // listed pixels with RT_SELCAP < c <= RT_MIDCAP: kept at the back of active[]
    int t_f, t_tile, t_len;              // current item (kept here across the sweep, which needs the registers)
    unsigned t_off;
    unsigned n_capped, n_big;
    int ncomp[RT_WARPS];                 // band items: faces of the warp's range that reach the band (-1: range not compacted)
};

// L2 residency of the fragment lists.  A list is written in P1, read once in P2 and dead afterwards; without
// hints its dirty lines are evicted to HBM by the streaming traffic around them and fetched back (measured:
// 1.39 GB of DRAM traffic per launch).  Stores carry an evict_last policy, and once a pixel is selected its
// lines (lists start on 128-byte boundaries) are dropped from L2 without write-back.
__device__ __forceinline__ unsigned long long l2_policy_evict_last() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void st_list_entry(uint2* p, unsigned a, unsigned b, unsigned long long pol) {
#ifdef RT_NO_L2HINTS
    (void)pol; *p = make_uint2(a, b);
#else
    asm volatile("st.global.L2::cache_hint.v2.b32 [%0], {%1, %2}, %3;" ::"l"(p), "r"(a), "r"(b), "l"(pol) : "memory");
#endif
}
// the slot a plane cursor points at: RT_LISTED | entry index (< 2^29) -> byte offset in 32 bits (the flag shifts out)
__device__ __forceinline__ uint2* list_slot(uint2* list, unsigned cursor) {
    return reinterpret_cast<uint2*>(reinterpret_cast<char*>(list) + (size_t)(cursor << 3));
}
__device__ __forceinline__ void l2_discard_line(const void* p) {
#if !defined(RT_NO_L2HINTS) && !defined(RT_NO_DISCARD)
    asm volatile("discard.global.L2 [%0], 128;" ::"l"(p) : "memory");
#else
    (void)p;
#endif
}
constexpr int RT_LIST_ALIGN = 16;               // entries per 128-byte line
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src_gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Face id of the slot-th (0-based) entry of the tile list whose rectangle covers pixel (lx, ly): list
// slots of a pixel are in tile-list order, so this recovers the face behind a slot (ties only).
__device__ unsigned rt_slot_to_fid(const uint4* __restrict__ pool, int len, int lx, int ly, int slot, int lane) {
    int run = 0;
    for (int base = 0; base < len; base += 32) {
        const int e = base + lane;
        bool cover = false;
        unsigned fid = 0u;
        if (e < len) {
            const uint4 en = pool[e];
            const int c0 = (int)(en.z & 0xffu), c1 = (int)((en.z >> 8) & 0xffu), r0 = (int)((en.z >> 16) & 0xffu), r1 = (int)(en.z >> 24);
            cover = lx >= c0 && lx <= c1 && ly >= r0 && ly <= r1;
            fid = en.x & 0xffffu;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, cover);
        const int n = __popc(bal);
        if (slot < run + n) return __shfl_sync(0xffffffffu, fid, (int)__fns(bal, 0u, slot - run + 1));
        run += n;
    }
    return 0xffffu;
}

// Product of (1 - p) over the K nearest valid fragments of one pixel's list L[0..c) (key, 1 - p); slots of
// rejected pairs carry key 0xffffffff and m = 1.  Slots are in face order, so "lower face id first" among
// equal depths is "lower slot first".  Returns the threshold: selected <=> key < tkey || (key == tkey &&
// slot <= tslot); tkey = 0xffffffff when every valid fragment is selected, tslot = -1 when no tie is cut.
__device__ __forceinline__ float rt_select(const uint2* __restrict__ L, const uint2* Ls /* shared-memory copy when c <= RT_SELCAP */,
                                           int c, int lane, unsigned* scratch /* >= RT_SELCAP words */,
                                           unsigned& tkey, int& tslot, bool& capped) {
    constexpr int NR = RT_SELCAP / 32;
    const unsigned ltmask = lanemask_lt();
    unsigned kr[NR];
    float mr[NR];
    tkey = 0xffffffffu; tslot = -1; capped = false;
    unsigned lo = 0xffffffffu, hi = 0u;
    int nv = 0;
    const bool small = c <= RT_SELCAP;
    {
        float pr = 1.f;
        if (small) {
#pragma unroll
            for (int r = 0; r < NR; ++r) {
                const int i = r * 32 + lane;
                kr[r] = 0xffffffffu; mr[r] = 1.f;
                if (i < c) { const uint2 e = Ls[i]; kr[r] = e.x; mr[r] = __uint_as_float(e.y); }
                pr *= mr[r];
                if (kr[r] != 0xffffffffu) { lo = min(lo, kr[r]); hi = max(hi, kr[r]); ++nv; }
            }
        } else {
            for (int i = lane; i < c; i += 32) {
                const uint2 e = L[i];
                pr *= __uint_as_float(e.y);
                if (e.x != 0xffffffffu) { lo = min(lo, e.x); hi = max(hi, e.x); ++nv; }
            }
        }
        nv = __reduce_add_sync(0xffffffffu, nv);
        if (nv <= RAST_K) return warp_prod(pr);
    }
    capped = true;
    lo = __reduce_min_sync(0xffffffffu, lo);
    hi = __reduce_max_sync(0xffffffffu, hi);
    bool exact = false;
    unsigned t = hi;
    int below = 0;                       // keys < lo
    int want = RAST_K;                   // rank looked for among the keys of the register phase
    if (!small) {
        // narrow the bracket on the whole list until at most RT_SELCAP keys remain in it
        int c_hi = nv;
        while (c_hi - below > RT_SELCAP && lo < hi) {
            const unsigned mid = lo + ((hi - lo) >> 1);
            int n = 0;
            for (int i = lane; i < c; i += 32) n += (L[i].x <= mid) ? 1 : 0;
            n = __reduce_add_sync(0xffffffffu, n);
            if (n == RAST_K) { t = mid; exact = true; break; }
            if (n > RAST_K) { hi = mid; c_hi = n; } else { lo = mid + 1; below = n; }
        }
        int m = 0;
        if (!exact && lo < hi) {
            for (int base = 0; base < c; base += 32) {
                const int i = base + lane;
                unsigned k = 0u;
                bool in = false;
                if (i < c) { k = L[i].x; in = (k >= lo && k <= hi); }
                const unsigned bal = __ballot_sync(0xffffffffu, in);
                if (in) scratch[m + __popc(bal & ltmask)] = k;
                m += __popc(bal);
            }
            __syncwarp();
            want = RAST_K - below;
        }
#pragma unroll
        for (int r = 0; r < NR; ++r) { const int i = r * 32 + lane; kr[r] = (i < m) ? scratch[i] : 0xffffffffu; }
        __syncwarp();
    }
    if (!exact) {
        while (lo < hi) {
            const unsigned mid = lo + ((hi - lo) >> 1);
            int n = 0;
#pragma unroll
            for (int r = 0; r < NR; ++r) n += (kr[r] <= mid) ? 1 : 0;
            n = __reduce_add_sync(0xffffffffu, n);
            if (n == want) { t = mid; exact = true; break; }
            if (n > want) hi = mid; else lo = mid + 1;
        }
    }
    int ts = -1;
    if (!exact) {
        t = lo;
        int clt = 0, cle = 0;
        const uint2* Lr = small ? Ls : L;
        for (int i = lane; i < c; i += 32) { const unsigned k = Lr[i].x; clt += (k < t); cle += (k <= t); }
        clt = __reduce_add_sync(0xffffffffu, clt);
        cle = __reduce_add_sync(0xffffffffu, cle);
        if (cle > RAST_K) {
            // ties on the depth at the cut: keep the first (K - clt) slots among key == t
            int need = RAST_K - clt;
            for (int base = 0; base < c; base += 32) {
                const int i = base + lane;
                const unsigned bal = __ballot_sync(0xffffffffu, i < c && Lr[i].x == t);
                const int n = __popc(bal);
                if (need <= n) { ts = base + (int)__fns(bal, 0u, need); break; }
                need -= n;
            }
        }
    }
    float prod = 1.f;
    if (small) {
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const int i = r * 32 + lane;
            // (in the small case kr[] still holds the pixel's own keys)
            if (kr[r] < t || (kr[r] == t && (ts < 0 || i <= ts))) prod *= mr[r];
        }
    } else {
        for (int i = lane; i < c; i += 32) {
            const uint2 e = L[i];
            if (e.x < t || (e.x == t && (ts < 0 || i <= ts))) prod *= __uint_as_float(e.y);
        }
    }
    tkey = t; tslot = ts;
    return warp_prod(prod);
}

// One prepared face (64-byte record in the warp's stage) against the pixels of its rectangle, 32 per step
// (rt_sweep_face2 below: 64 per step).
// SKIPS: some pixels of the tile are not handled in this pass (multi-pass tiles only).
// Lanes walk the rectangle's pixels (or pixel pairs) in row-major order, 32 per step: position (rr, cc) of a lane's
// item advances by 32 = sq * wd + sr per step, with one carry (no division in the loop).
struct RectWalk {
    int rr, cc, sq, sr;
    __device__ __forceinline__ RectWalk(int wd, int lane) {
        float inv_w;      // MUFU reciprocal: (i + 0.5) / wd is at least 0.5 / wd away from an integer; exact for wd <= 1024
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv_w) : "f"((float)wd));
        rr = (int)(((float)lane + 0.5f) * inv_w);
        cc = lane - rr * wd;
        sq = (int)(32.5f * inv_w);
        sr = 32 - sq * wd;
    }
    __device__ __forceinline__ void next(int wd) {
        rr += sq; cc += sr;
        if (cc >= wd) { cc -= wd; ++rr; }
    }
};

template <bool SKIPS, bool REGULAR>
__device__ __forceinline__ void rt_sweep_face1(const FaceSetup& fs, unsigned* __restrict__ plane, uint2* __restrict__ list,
                                               int lane, int x0, int y0, float inv_s, int c0, int c1, int r0, int r1,
                                               unsigned long long pol) {
    const int wd = c1 - c0 + 1, npx = wd * (r1 - r0 + 1);
    RectWalk wk(wd, lane);
    for (int i = lane; i < npx; i += 32, wk.next(wd)) {
        const int lx = c0 + wk.cc, ly = r0 + wk.rr;
        const int idx = ly * RT_PITCH + lx;
        const unsigned v = plane[idx];
        if (SKIPS && v == RT_SKIP) continue;
        float sd, pz, mv = 1.f;
        const bool ok = frag_setup_forward<REGULAR>(fs, pix_to_ndc(x0 + lx, inv_s), pix_to_ndc(y0 + ly, inv_s), sd, pz);
        if (ok) { float pp; frag_prob(sd, pp, mv); }
        if (v & RT_LISTED) {
            plane[idx] = v + 1u;
            st_list_entry(list_slot(list, v), ok ? __float_as_uint(pz + 0.f) : 0xffffffffu, __float_as_uint(mv), pol);
        } else if (ok) {
            plane[idx] = __float_as_uint(__uint_as_float(v) * mv);
        }
    }
    __syncwarp();        // the next face's lanes may touch the same pixels
}

// The same sweep two horizontally adjacent pixels per lane (packed FP32, face_eval2): 64 pixels per step.  Used for
// rectangles of more than 32 pixels of faces without degenerate edges; results are bit-identical to rt_sweep_face.
template <bool SKIPS>
__device__ __forceinline__ void rt_sweep_face2(const FaceSetup& fs, unsigned* __restrict__ plane, uint2* __restrict__ list,
                                               int lane, int x0, int y0, float inv_s, int c0, int c1, int r0, int r1,
                                               unsigned long long pol) {
    const int wp = (c1 - c0 + 2) >> 1, npairs = wp * (r1 - r0 + 1);
    RectWalk wk(wp, lane);
    for (int j = lane; j < npairs; j += 32, wk.next(wp)) {
        const int lx = c0 + 2 * wk.cc, ly = r0 + wk.rr;
        const int idx = ly * RT_PITCH + lx;
        const bool has1 = lx < c1;
        unsigned v[2];
        v[0] = plane[idx];
        v[1] = plane[idx + 1];           // (column 32 of the pitch-33 plane exists; never written when !has1)
        bool act[2] = {true, has1};
        if (SKIPS) { act[0] = v[0] != RT_SKIP; act[1] = has1 && v[1] != RT_SKIP; if (!act[0] && !act[1]) continue; }
        const float t0 = ffma(2.f, (float)(x0 + lx), 1.f);
        const f32x2 px = f2_fma(f2_pack(t0, t0 + 2.f), f2_bc(-inv_s), f2_bc(1.f));       // pix_to_ndc of both columns
        bool ok[2];
        Fragment2 fr;
        face_eval2<false>(fs, px, pix_to_ndc(y0 + ly, inv_s), ok, fr);
        float pp[2], mv[2];
        frag_prob2(fr.sd, pp, mv);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            if (!act[k]) continue;
            if (v[k] & RT_LISTED) {
                plane[idx + k] = v[k] + 1u;
                st_list_entry(list_slot(list, v[k]), ok[k] ? __float_as_uint(fr.pz[k] + 0.f) : 0xffffffffu,
                              __float_as_uint(ok[k] ? mv[k] : 1.f), pol);
            } else if (ok[k]) {
                plane[idx + k] = __float_as_uint(__uint_as_float(v[k]) * mv[k]);
            }
        }
    }
    __syncwarp();
}

#ifndef RT_PACK_MIN
#define RT_PACK_MIN 33          // rectangles of at least this many pixels are swept two pixels per lane
#endif
template <bool SKIPS>
__device__ __forceinline__ void rt_sweep_face(const float4* __restrict__ rec, unsigned* __restrict__ plane, uint2* __restrict__ list,
                                              int lane, int x0, int y0, float inv_s, int b0, int b1, unsigned long long pol) {
    const float4 q3 = rec[3];
    const unsigned rect = __float_as_uint(q3.z);
    const int c0 = (int)(rect & 0xffu), c1 = (int)((rect >> 8) & 0xffu);
    const int r0 = max((int)((rect >> 16) & 0xffu), b0), r1 = min((int)(rect >> 24), b1 - 1);     // rows of this item only
    if (r0 > r1) return;
    const float4 q0 = rec[0], q1 = rec[1], q2 = rec[2];
    FaceSetup fs;
    fs.x0 = q0.x; fs.y0 = q0.y; fs.x1 = q0.z; fs.y1 = q0.w; fs.x2 = q1.x; fs.y2 = q1.y;
    fs.z0 = q1.z; fs.z1 = q1.w; fs.z2 = q2.x; fs.rden = q2.y;
    fs.rl01 = q2.z; fs.rl02 = q2.w; fs.rl12 = q3.x;
    fs.e01x = fsub(fs.x1, fs.x0); fs.e01y = fsub(fs.y1, fs.y0);
    fs.e02x = fsub(fs.x2, fs.x0); fs.e02y = fsub(fs.y2, fs.y0);
    fs.e12x = fsub(fs.x2, fs.x1); fs.e12y = fsub(fs.y2, fs.y1);
    const int npx = (c1 - c0 + 1) * (r1 - r0 + 1);
    const bool regular = (fs.rl01 != 0.f) && (fs.rl02 != 0.f) && (fs.rl12 != 0.f);       // no degenerate edge (warp-uniform)
    if (!regular) rt_sweep_face1<SKIPS, false>(fs, plane, list, lane, x0, y0, inv_s, c0, c1, r0, r1, pol);
    else if (npx >= RT_PACK_MIN) rt_sweep_face2<SKIPS>(fs, plane, list, lane, x0, y0, inv_s, c0, c1, r0, r1, pol);
    else rt_sweep_face1<SKIPS, true>(fs, plane, list, lane, x0, y0, inv_s, c0, c1, r0, r1, pol);
}

__global__ void __launch_bounds__(RT_THREADS, RT_CTAS_PER_SM)
raster_tile_forward_kernel(ModelDev m, Workspace w, TileScratch ts, int frame0, int n_frames, Weights wt, float* alpha_out) {
    grid_dep_wait();
    extern __shared__ __align__(128) unsigned char smem_raw[];
    RtSmem& sm = *reinterpret_cast<RtSmem*>(smem_raw);
    int lane;                            // (asm volatile: kept in a register instead of being re-derived from S2R in the inner loops)
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane));
    const int wid = threadIdx.x >> 5;
    unsigned* plane = sm.plane[wid];
    if (lane == 0) { mbar_init(&sm.bar[wid][0], 1); mbar_init(&sm.bar[wid][1], 1); mbar_fence_init(); }
    if (threadIdx.x == 0) { sm.n_capped = 0u; sm.n_big = 0u; }
    unsigned phase = 0;                  // bit b: parity the warp waits for next on its stage buffer b

    // Values that are only needed again after the sweep live in sm.t across it (the sweep needs the registers).
    // The hand-out counter is a global atomic (~1 us round trip with the whole CTA waiting) and the item's record one more
    // dependent load.  (Drawing the next item ahead -- at the start of an item in round 1, after its sweep in round 2 -- was
    // measured worse both times, +2 ... +7 %: items reserved by busy CTAs are missing from the tail of the launch.)
    // Items sit in RT_ITEM_BINS size classes, class 0 = largest (emit_items in frame_front): the k-th draw takes the k-th item
    // of the classes laid end to end.
    if (threadIdx.x < RT_ITEM_BINS) {
        unsigned end = 0u;
        for (int b = 0; b <= (int)threadIdx.x; ++b) end += ts.bin_count[b];
        sm.bin_end[threadIdx.x] = end;
    }
    for (;;) {
        __syncthreads();                 // the previous item's shared state is no longer read (first round: bin_end is written)
        if (threadIdx.x == 0) {
            const unsigned k = atomicAdd(ts.item_next, 1u);
            int b = 0;
            while (b < RT_ITEM_BINS && k >= sm.bin_end[b]) ++b;
            sm.item = (b < RT_ITEM_BINS) ? (long long)b * ts.bin_cap + (k - (b ? sm.bin_end[b - 1] : 0u)) : -1ll;
        }
        __syncthreads();
        const long long item = sm.item;
        if (item < 0) break;
        // item = one tile of one frame, or one band of rows of a tile with a long list: code, list offset, length
        const uint4 rec = ts.items[item];
        const unsigned code = rec.x;
        const int f = (int)(code >> 15), fr = frame0 + f, tile = (int)((code >> 5) & 0x3ffu);
        const int T = w.tiles_x * w.tiles_y;
        const int bh = TILE_H >> (code & 3u), b0 = (int)((code >> 2) & 7u) * bh, b1 = b0 + bh;
        const unsigned off = rec.y;
        const int len = (int)rec.z;
        const int x0 = (tile % w.tiles_x) * TILE_W, y0 = (tile / w.tiles_x) * TILE_H;
        if (threadIdx.x == 0) { sm.t_f = f; sm.t_tile = tile; sm.t_len = len; sm.t_off = off; }
        if (len == 0) {
            // no face reaches the tile: alpha = 0, |alpha - T| = T; pix is never read here
            const size_t slot0 = ((size_t)fr * T + (size_t)tile) * (REGIONS_PER_TILE * REGION_H);
            if (threadIdx.x < REGIONS_PER_TILE * REGION_H) w.region_l1[slot0 + threadIdx.x] = w.region_tsum[slot0 + threadIdx.x];
            if (alpha_out) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int x = x0 + lane, y = y0 + wid + RT_WARPS * k;
                    if (x < w.S && y < w.S) alpha_out[((size_t)f * w.S + y) * w.S + x] = 0.f;
                }
            }
            continue;
        }
        const uint4* pool = w.tile_pool + (size_t)fr * w.pool_cap + off;
        // Split the list into RT_WARPS contiguous ranges of about equal cost (pairs + a per-face overhead),
        // at RT_BLK granularity: the warps meet at CTA barriers, the slowest one sets the pace.
        int lo, hi;
        const int nchunk = (len + RT_BLK - 1) / RT_BLK;
        if (nchunk <= RT_MAXCHUNK) {
            for (int base = 0; base < nchunk * RT_BLK; base += RT_THREADS) {
                const int e = base + (int)threadIdx.x;
                unsigned cost = 0u;
                if (e < len) {
                    const unsigned rect = pool[e].z;
                    const int rows = min((int)(rect >> 24), b1 - 1) - max((int)((rect >> 16) & 0xffu), b0) + 1;
                    const unsigned npx = (((rect >> 8) & 0xffu) - (rect & 0xffu) + 1u) * (unsigned)max(rows, 0);
                    // the sweep takes whole 32-lane steps: one pixel per lane, or -- rectangles of RT_PACK_MIN pixels and
                    // more -- a pixel pair per lane at about 1.3x the cost of a step
#ifdef RT_OLD_COST
                    if (false) {
#else
                    if (npx >= (unsigned)RT_PACK_MIN) {
#endif
                        const unsigned wdp = (((rect >> 8) & 0xffu) - (rect & 0xffu) + 2u) >> 1;
                        cost = ((wdp * (unsigned)rows + 31u) >> 5) * 42u + RT_FACE_COST;
                    } else {
                        cost = npx ? ((npx + 31u) & ~31u) + RT_FACE_COST : 2u;
                    }
                }
#pragma unroll
                for (int o = RT_BLK / 2; o > 0; o >>= 1) cost += __shfl_xor_sync(0xffffffffu, cost, o);
                if ((lane & (RT_BLK - 1)) == 0 && e < nchunk * RT_BLK) sm.chunk[e / RT_BLK] = cost;
            }
            __syncthreads();
            if (wid == 0) {
                constexpr int PER = RT_MAXCHUNK / 32;
                unsigned loc[PER];
                unsigned sum = 0u;
#pragma unroll
                for (int k = 0; k < PER; ++k) { const int ch = lane * PER + k; loc[k] = (ch < nchunk) ? sm.chunk[ch] : 0u; sum += loc[k]; }
                unsigned incl = sum;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { const unsigned y = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += y; }
                unsigned run = incl - sum;
#pragma unroll
                for (int k = 0; k < PER; ++k) { const int ch = lane * PER + k; if (ch < nchunk) sm.chunk[ch] = run; run += loc[k]; }
                if (lane == 31) sm.cost_total = incl;
            }
            __syncthreads();
            const unsigned long long tot = sm.cost_total;
            const unsigned t_lo = (unsigned)((tot * (unsigned)wid) / RT_WARPS), t_hi = (unsigned)((tot * (unsigned)(wid + 1)) / RT_WARPS);
            int n_lo = 0, n_hi = 0;          // chunks that start before the target: the exclusive prefix is non-decreasing
            for (int ch = lane; ch < nchunk; ch += 32) { const unsigned v = sm.chunk[ch]; n_lo += (v < t_lo) ? 1 : 0; n_hi += (v < t_hi) ? 1 : 0; }
            n_lo = __reduce_add_sync(0xffffffffu, n_lo);
            n_hi = __reduce_add_sync(0xffffffffu, n_hi);
            if (wid == RT_WARPS - 1) n_hi = nchunk;
            lo = min(n_lo * RT_BLK, len); hi = min(n_hi * RT_BLK, len);
        } else {
            const int per = ((nchunk + RT_WARPS - 1) / RT_WARPS) * RT_BLK;
            lo = min(wid * per, len); hi = min(lo + per, len);
        }

        for (unsigned round = 0;; ++round) {
            // ---- P0: candidates per (warp, pixel) = integral of the rectangles' corner grid
            // (only the rows of the item's band, plus the closing corner row, are ever read back)
            for (int i = b0 * RT_PITCH + lane; i < (b1 + 1) * RT_PITCH + 1; i += 32) plane[i] = 0u;
            __syncwarp();
            // A band item only sweeps the faces of the tile list that reach its rows: their positions are compacted here
            // (in list order) and P1 gathers those records instead of streaming the whole list -- a heavy tile cut into
            // 8 bands would otherwise visit every face 8 times.
#ifdef RT_NO_BAND_COMPACT
            const bool compact = false;
#else
            const bool compact = (bh < TILE_H) && (hi - lo <= RT_BAND_MAX);
#endif
            unsigned short* bidx = ts.band_idx + ((size_t)blockIdx.x * RT_WARPS + wid) * RT_BAND_MAX;
            {
                int nc = 0;
                for (int base = lo; base < hi; base += 32) {
                    const int e = base + lane;
                    bool ov = false;
                    if (e < hi) {
                        const unsigned rect = pool[e].z;
                        const int c0 = (int)(rect & 0xffu), c1 = (int)((rect >> 8) & 0xffu);
                        const int r0 = max((int)((rect >> 16) & 0xffu), b0), r1 = min((int)(rect >> 24), b1 - 1);
                        ov = r0 <= r1;
                        if (ov) {
                            atomicAdd(&plane[r0 * RT_PITCH + c0], 1u);
                            atomicAdd(&plane[r0 * RT_PITCH + c1 + 1], 0xffffffffu);
                            atomicAdd(&plane[(r1 + 1) * RT_PITCH + c0], 0xffffffffu);
                            atomicAdd(&plane[(r1 + 1) * RT_PITCH + c1 + 1], 1u);
                        }
                    }
                    if (compact) {
                        const unsigned bal = __ballot_sync(0xffffffffu, ov);
                        if (ov) bidx[nc + __popc(bal & lanemask_lt())] = (unsigned short)(e - lo);
                        nc += __popc(bal);
                    }
                }
                if (lane == 0) sm.ncomp[wid] = compact ? nc : -1;
            }
            __syncwarp();
            {
                unsigned run = 0u;
#pragma unroll 4
                for (int r = b0; r < b1; ++r) { run += plane[r * RT_PITCH + lane]; plane[r * RT_PITCH + lane] = run; }
            }
            __syncwarp();
            {
                unsigned run = 0u;
                if (lane >= b0 && lane < b1) {
#pragma unroll 8
                    for (int c = 0; c < TILE_W; ++c) { run += plane[lane * RT_PITCH + c]; plane[lane * RT_PITCH + c] = run; }
                }
            }
            __syncthreads();

            // ---- P0b: classify the pixels (thread: column = lane, rows wid + 8k), lay out the lists
            {
                const unsigned cap = (unsigned)ts.list_cap;
                unsigned cnt4[4];
                unsigned mysum = 0u;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int row = wid + RT_WARPS * k;
                    const int idx = row * RT_PITCH + lane;
                    unsigned c = 0u;
                    if (row >= b0 && row < b1) {             // (rows outside the item's band: nothing to do)
#pragma unroll
                        for (int q = 0; q < RT_WARPS; ++q) c += sm.plane[q][idx];
                    }
                    cnt4[k] = c;
                    if (c > (unsigned)RAST_K) mysum += (c + RT_LIST_ALIGN - 1) & ~(unsigned)(RT_LIST_ALIGN - 1);     // lists start on 128-byte lines
                }
                unsigned incl = mysum;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { const unsigned y = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += y; }
                if (lane == 31) sm.warp_tot[wid] = incl;
                if (threadIdx.x == 0) { sm.n_active = 0u; sm.p2_next = 0u; sm.n_mid = 0u; }
                __syncthreads();
                unsigned base = 0u, total = 0u;
#pragma unroll
                for (int q = 0; q < RT_WARPS; ++q) { const unsigned v = sm.warp_tot[q]; if (q < wid) base += v; total += v; }
                if (threadIdx.x == 0) sm.total = total;
                unsigned run = base + incl - mysum;
                const unsigned win_lo = round * cap;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int row = wid + RT_WARPS * k;
                    if (row < b0 || row >= b1) continue;
                    const int idx = row * RT_PITCH + lane, px = row * TILE_W + lane;
                    const unsigned c = cnt4[k];
                    unsigned char cls;
                    if (c > (unsigned)RAST_K) {
                        const unsigned o = run;
                        run += (c + RT_LIST_ALIGN - 1) & ~(unsigned)(RT_LIST_ALIGN - 1);
                        const bool act = (o >= win_lo) && (o - win_lo < cap);
                        // (cursors address the whole scratch, not this CTA's part: the sweep then forms a store address from
                        //  the kernel parameter alone; n_ctas * list_stride < 2^29 is checked at create)
                        unsigned cur = RT_LISTED | (blockIdx.x * (unsigned)ts.list_stride + (o - win_lo));
#pragma unroll
                        for (int q = 0; q < RT_WARPS; ++q) {
                            const unsigned cw = sm.plane[q][idx];
                            sm.plane[q][idx] = act ? cur : RT_SKIP;
                            cur += cw;
                        }
                        if (act) {
                            if (c > (unsigned)RT_SELCAP && c <= (unsigned)RT_MIDCAP) {
                                const unsigned a = atomicAdd(&sm.n_mid, 1u);
                                sm.active[RT_PIX - 1 - a] = (unsigned short)px;
                            } else {
                                const unsigned a = atomicAdd(&sm.n_active, 1u);
                                sm.active[a] = (unsigned short)px;
                            }
                            sm.list_off[px] = o - win_lo;
                            sm.list_cnt[px] = (unsigned short)c;
                        }
                        cls = act ? RT_CLS_LISTED : RT_CLS_IDLE;
                    } else {
                        const unsigned v = (round == 0u) ? 0x3f800000u : RT_SKIP;
#pragma unroll
                        for (int q = 0; q < RT_WARPS; ++q) sm.plane[q][idx] = v;
                        cls = (round == 0u) ? RT_CLS_DIRECT : RT_CLS_IDLE;
                    }
                    sm.cls[px] = cls;
                }
            }
            __syncthreads();

            // ---- P1: sweep this warp's faces (TMA double buffer of RT_BLK prepared faces)
            {
                const float4* recs = w.tile_rec + ((size_t)fr * w.pool_cap + off) * 4;
                uint2* list = ts.list;             // cursors carry the CTA's offset
                const float inv_s = 1.f / (float)w.S;
                const int ncomp = sm.ncomp[wid];
                const bool gather = ncomp >= 0;          // band item: the compacted faces are gathered with cp.async
                const int nface = gather ? ncomp : hi - lo;
                const int nblk = (nface + RT_BLK - 1) / RT_BLK;
                const bool skips = (round != 0u) || (sm.total > (unsigned)ts.list_cap);       // otherwise no plane holds RT_SKIP
                const unsigned long long pol = l2_policy_evict_last();
                const unsigned short* bidx = ts.band_idx + ((size_t)blockIdx.x * RT_WARPS + wid) * RT_BAND_MAX;
                // block b of RT_BLK prepared faces -> stage buffer b & 1: one bulk copy (contiguous part of the list), or
                // two 16-byte cp.async per lane (lane pair 2j, 2j + 1 fetches the two halves of compacted face j)
                auto fetch = [&](int b) {
                    if (gather) {
                        const int k = b * RT_BLK + (lane >> 1);
                        if (k < nface) {
                            const float4* src = recs + ((size_t)lo + bidx[k]) * 4 + (lane & 1) * 2;
                            float4* dst = sm.stage[wid][b & 1] + (lane >> 1) * 4 + (lane & 1) * 2;
                            cp_async16(dst, src);
                            cp_async16(dst + 1, src + 1);
                        }
                        cp_async_commit();
                    } else if (lane == 0) {
                        const int e0 = lo + b * RT_BLK;
                        const unsigned bytes = (unsigned)min(RT_BLK, hi - e0) * 64u;
                        mbar_expect_tx(&sm.bar[wid][b & 1], bytes);
                        tma_load_1d(sm.stage[wid][b & 1], recs + (size_t)e0 * 4, bytes, &sm.bar[wid][b & 1]);
                    }
                };
                if (nblk > 0) {
                    // the stage doubles as P2's scratch (generic-proxy writes): order them before the bulk copies
                    if (!gather && lane == 0) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    fetch(0);
                }
                for (int b = 0; b < nblk; ++b) {
                    if (b + 1 < nblk) fetch(b + 1);
                    else if (gather) cp_async_commit();
                    if (gather) {
                        cp_async_wait<1>();              // everything but the group just committed has landed
                        __syncwarp();
                    } else {
                        mbar_wait(&sm.bar[wid][b & 1], (phase >> (b & 1)) & 1u);
                        phase ^= 1u << (b & 1);
                    }
                    const int nrec = min(RT_BLK, nface - b * RT_BLK);
                    const float4* st = sm.stage[wid][b & 1];
                    if (skips) { for (int j = 0; j < nrec; ++j) rt_sweep_face<true>(st + j * 4, plane, list, lane, x0, y0, inv_s, b0, b1, pol); }
                    else       { for (int j = 0; j < nrec; ++j) rt_sweep_face<false>(st + j * 4, plane, list, lane, x0, y0, inv_s, b0, b1, pol); }
                }
                if (gather) cp_async_wait<0>();
            }
            __syncthreads();

            // ---- P2a: the direct pixels' products leave the planes (fixed warp order) ...
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int row = wid + RT_WARPS * k;
                const int idx = row * RT_PITCH + lane;
                if (row < b0 || row >= b1 || sm.cls[row * TILE_W + lane] != RT_CLS_DIRECT) continue;
                float P = __uint_as_float(sm.plane[0][idx]);
#pragma unroll
                for (int q = 1; q < RT_WARPS; ++q) P *= __uint_as_float(sm.plane[q][idx]);
                sm.pl[row * TILE_W + lane] = P;
            }
            __syncthreads();

            // ---- P2: ... and every warp's plane becomes staging space for the listed pixels of this pass: one warp per
            //      pixel.  Lists of up to RT_SELCAP entries use half a plane each, the next pixel's list streaming in
            //      (cp.async) while this one is selected; lists of up to RT_MIDCAP entries take the whole plane and are
            //      narrowed there; longer ones are narrowed in global memory.
            {
                const unsigned n_short = sm.n_active, na = n_short + sm.n_mid;        // mid-length lists come last
                const uint2* list = ts.list + (size_t)blockIdx.x * ts.list_stride;
                unsigned* scratch = reinterpret_cast<unsigned*>(sm.stage[wid][0]);     // the TMA stage is idle here
                uint2* stg = reinterpret_cast<uint2*>(plane);                          // 2 x RT_SELCAP entries
                auto pixel_of = [&](unsigned a) { return (int)sm.active[a < n_short ? a : (RT_PIX - 1) - (a - n_short)]; };
                auto stage_px = [&](unsigned a, uint2* dst, int cap) {
                    const int px = pixel_of(a);
                    const int c = (int)sm.list_cnt[px];
                    if (c <= cap) {
                        const uint2* src = list + sm.list_off[px];
                        for (int ch = lane; ch * 2 < c; ch += 32) cp_async16(dst + ch * 2, src + ch * 2);
                    }
                };
                // pixels are drawn from a shared counter (lists differ a lot in length; the results do not depend on
                // which warp takes which pixel), one ahead so that the next list can stream in
                auto draw = [&]() { unsigned a = 0u; if (lane == 0) a = atomicAdd(&sm.p2_next, 1u); return __shfl_sync(0xffffffffu, a, 0); };
                int buf = 0;
                unsigned a = draw();
                if (a < n_short) stage_px(a, stg, RT_SELCAP);
                cp_async_commit();
                while (a < na) {
                    const unsigned a_next = draw();
                    const bool mid = a >= n_short;
                    if (mid) {
                        stage_px(a, stg, RT_MIDCAP);         // (nothing else is in flight: mid-length pixels are never prefetched)
                        cp_async_commit();
                        cp_async_wait<0>();
                    } else {
                        if (a_next < n_short) stage_px(a_next, stg + (buf ^ 1) * RT_SELCAP, RT_SELCAP);
                        cp_async_commit();
                        cp_async_wait<1>();              // everything but the group just committed has landed
                    }
                    __syncwarp();
                    const int px = pixel_of(a);
                    const int c = (int)sm.list_cnt[px];
                    const uint2* src = list + sm.list_off[px];
                    unsigned tk, tf = 0xffffu;
                    int tslot;
                    bool capped;
                    const float P = rt_select(mid ? stg : src, mid ? stg : stg + buf * RT_SELCAP, c, lane, scratch, tk, tslot, capped);
                    for (int ln = lane * RT_LIST_ALIGN; ln < c; ln += 32 * RT_LIST_ALIGN) l2_discard_line(src + ln);   // dead from here on
                    if (tslot >= 0)
                        tf = rt_slot_to_fid(w.tile_pool + (size_t)(frame0 + sm.t_f) * w.pool_cap + sm.t_off, sm.t_len, px % TILE_W, px / TILE_W, tslot, lane);
                    __syncwarp();                    // every lane is done with this buffer (it is refilled two pixels on)
                    if (lane == 0) {
                        sm.pl[px] = P;
                        sm.list_off[px] = tk;
                        sm.list_cnt[px] = (unsigned short)tf;
                        if (capped) atomicAdd(&sm.n_capped, 1u);
                        if (c > RT_SELCAP) atomicAdd(&sm.n_big, 1u);
                    }
                    buf ^= 1;
                    a = a_next;
                }
                cp_async_wait<0>();
            }
            __syncthreads();

            // ---- P3: finish the pixels of this pass
            {
                const int S = w.S, f2 = sm.t_f, fr2 = frame0 + f2, tile2 = sm.t_tile;
                const int x = (tile2 % w.tiles_x) * TILE_W + lane;
                const float inv_s = 1.f / (float)S;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int row = wid + RT_WARPS * k;
                    const int px = row * TILE_W + lane;
                    if (row < b0 || row >= b1) continue;
                    const unsigned cls = sm.cls[px];
                    if (cls == RT_CLS_IDLE) continue;
                    const float P = sm.pl[px];
                    unsigned tk = 0xffffffffu, tf = 0xffffu;
                    if (cls == RT_CLS_LISTED) { tk = sm.list_off[px]; tf = sm.list_cnt[px]; }
                    const int y = (tile2 / w.tiles_x) * TILE_H + row;
                    float l1 = 0.f;
                    if (x < S && y < S) {
                        const size_t pi = ((size_t)fr2 * S + y) * S + x;
                        const float alpha = 1.f - P;
                        const float d = alpha - (float)w.sil[pi];
                        l1 = fabsf(d);
                        float coef = 0.f;
                        if (P < 1.f && P >= P_SKIP && d != 0.f) {        // P < 1 <=> the pixel has fragments
                            const float ga = wt.sil * w.inv_window[fr2] * inv_s * inv_s * (d > 0.f ? 1.f : -1.f);
                            coef = ga * P * (1.f / RAST_SIGMA);
                        }
                        w.pix[pi] = make_uint2(__float_as_uint(coef), tk);
                        if (tk != 0xffffffffu) w.pix_tfid[pi] = (unsigned short)tf;
                        if (alpha_out) alpha_out[((size_t)f2 * S + y) * S + x] = alpha;
                    }
                    sm.l1[px] = l1;
                }
            }
            if (sm.total <= (round + 1u) * (unsigned)ts.list_cap) break;      // every listed pixel started inside a window already done
            __syncthreads();                             // planes are rebuilt by the next pass
        }
        __syncthreads();
        // per region row (8 pixels), fixed order
        if (threadIdx.x < REGIONS_PER_TILE * REGION_H) {
            const int sub = threadIdx.x / REGION_H, row = threadIdx.x % REGION_H;
            const int lx0 = (sub % (TILE_W / REGION_W)) * REGION_W, ly0 = (sub / (TILE_W / REGION_W)) * REGION_H;
            const float* p = sm.l1 + (ly0 + row) * TILE_W + lx0;
            const size_t slot0 = ((size_t)(frame0 + sm.t_f) * (w.tiles_x * w.tiles_y) + (size_t)sm.t_tile) * (REGIONS_PER_TILE * REGION_H);
            if (ly0 >= b0 && ly0 < b1)       // region rows of this item's band
                w.region_l1[slot0 + threadIdx.x] = ((p[0] + p[1]) + (p[2] + p[3])) + ((p[4] + p[5]) + (p[6] + p[7]));
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (sm.n_capped | sm.n_big) {
            atomicAdd(w.counters + 0, (unsigned long long)sm.n_capped);
            atomicAdd(w.counters + 1, (unsigned long long)sm.n_big);
        }
        // the last CTA to finish clears the hand-out state for the next launch and passes the pair total on
        __threadfence();
        if (atomicAdd(ts.exit_ticket, 1u) == gridDim.x - 1u) {
            __threadfence();
            for (int b = 0; b < RT_ITEM_BINS; ++b) ts.bin_count[b] = 0u;
            *ts.item_next = 0u;
            *ts.exit_ticket = 0u;
            *ts.prev_total = *(volatile unsigned long long*)ts.total_cost;
            *ts.total_cost = 0ull;
        }
    }
}

size_t raster_tile_smem_bytes() { return sizeof(RtSmem); }

void launch_raster_tile_forward(const ModelDev& m, const Workspace& w, const TileScratch& ts, int frame0, int n, Weights wt,
                                float* alpha_out, int n_ctas, cudaStream_t st) {
    // (the items were emitted by frame_front, frame by frame)
    const long long tiles = (long long)n * w.tiles_x * w.tiles_y;
    const int grid = (int)(tiles < n_ctas ? tiles : n_ctas);
    launch_pdl(raster_tile_forward_kernel, dim3(grid), dim3(RT_THREADS), raster_tile_smem_bytes(), st, m, w, ts, frame0, n, wt, alpha_out);
}

}  // namespace smf
