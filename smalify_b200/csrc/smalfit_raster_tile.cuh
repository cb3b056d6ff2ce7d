// smalfit_raster_tile.cuh -- forward soft rasteriser, tile formulation (included by smalfit_kernels.cu).
//
// Semantics: PyTorch3D 0.2.5 rasterize_meshes (faces_per_pixel = 100, blur 9.21e-4, no culling) +
// sigmoid_alpha_blend as configured at smal_fitter/p3d_renderer.py:26-39,66, fused with the L1
// silhouette term of smal_fitter/smal_fitter.py:172-173.  Same outputs as raster_forward_kernel:
// per pixel (coef, depth threshold, tie face id) for raster_backward, per region row sum|alpha - T|.
//
// Work item = one 32x32-pixel tile of one frame (bin_faces lists, heaviest tiles of every frame
// first), pulled by persistent CTAs of 8 warps, 3 CTAs per SM.  The tile's face list is split into
// 8 contiguous ranges, one per warp, and every warp is *face-parallel*: it streams its prepared
// faces (64-byte records written by bin_faces) through a double-buffered shared-memory stage with
// 1-D TMA bulk copies, and for each face its 32 lanes sweep the face's pixel rectangle.  No pair is
// evaluated twice and nothing is gathered: the per-pair cost is the fragment arithmetic itself.
//
//   P0  box counts: each warp adds its faces' rectangles into its own 33x33 corner grid (native
//       32-bit shared-memory atomics) and integrates it -> candidates per (warp, pixel).
//   P0b per pixel: c = sum over warps.  c <= K: every fragment is selected, the pixel only needs the
//       product of (1 - p): the planes are set to 1.0f.  c > K ("listed"): the pixel gets c slots in
//       the CTA's fragment list (global scratch, L2 resident); each warp's plane holds its write
//       cursor = list offset + candidates of the warps before it, so slots are in face order and
//       the list is identical from run to run.  Tiles whose lists exceed the scratch are done in
//       several passes over disjoint pixel sets.
//   P1  sweep: direct pixels multiply into the warp's plane (plain LDS/FMUL/STS: lanes of one face
//       touch distinct pixels), listed pixels store (depth key, 1 - p, face id) at their cursor.
//   P2  per listed pixel, one warp: keys in registers (8 per lane), exact K-th order statistic of
//       (depth, face id) by bisection on the key bits, product over the selected set.
//   P3  per pixel: alpha, |alpha - T|, coef = dL/dalpha * P / sigma; region row sums in fixed order.
// Every reduction has a fixed order: results are run-to-run deterministic.
#pragma once

namespace smf {

constexpr unsigned RT_SKIP = 0xffffffffu;       // plane value: pixel not handled in this pass
constexpr unsigned RT_LISTED = 0x80000000u;     // plane value: write cursor of a listed pixel (low 31 bits)
constexpr int RT_SELCAP = 256;                  // keys of one pixel held in registers
constexpr int RT_PIX = TILE_W * TILE_H;

struct RtSmem {
    unsigned plane[RT_WARPS][RT_PLANE];
    float4 stage[RT_WARPS][2][RT_BLK * 4];
    float l1[RT_PIX];
    unsigned list_off[RT_PIX];           // listed pixels: list offset, then (after P2) the depth threshold
    unsigned short list_cnt[RT_PIX];     // listed pixels: slots, then (after P2) the tie face id
    unsigned short active[RT_PIX];
    unsigned long long bar[RT_WARPS][2];
    unsigned warp_tot[RT_WARPS];
    int item;
    unsigned n_active;
};

// Product of (1 - p) over the K nearest valid fragments of one pixel's list L[0..c) (slots of rejected
// pairs carry key 0xffffffff and m = 1).  Threshold (tkey, tfid): selected <=> key < tkey ||
// (key == tkey && fid <= tfid); (0xffffffff, 0xffff) when every valid fragment is selected.
__device__ float rt_select(const uint4* __restrict__ L, int c, int lane, unsigned* scratch /* >= RT_SELCAP words */,
                           unsigned& tkey, unsigned& tfid, bool& capped) {
    constexpr int NR = RT_SELCAP / 32;
    const unsigned ltmask = lanemask_lt();
    unsigned kr[NR];
    tkey = 0xffffffffu; tfid = 0xffffu; capped = false;
    unsigned lo = 0xffffffffu, hi = 0u;
    int nv = 0;
    if (c <= RT_SELCAP) {
        float pr = 1.f;
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const int i = r * 32 + lane;
            kr[r] = 0xffffffffu;
            if (i < c) { const uint4 e = L[i]; kr[r] = e.x; pr *= __uint_as_float(e.y); }
            if (kr[r] != 0xffffffffu) { lo = min(lo, kr[r]); hi = max(hi, kr[r]); ++nv; }
        }
        nv = __reduce_add_sync(0xffffffffu, nv);
        if (nv <= RAST_K) return warp_prod(pr);
    } else {
        float pr = 1.f;
        for (int i = lane; i < c; i += 32) {
            const uint4 e = L[i];
            pr *= __uint_as_float(e.y);
            if (e.x != 0xffffffffu) { lo = min(lo, e.x); hi = max(hi, e.x); ++nv; }
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
    if (c > RT_SELCAP) {
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
    unsigned tf = 0xffffu;
    if (!exact) {
        t = lo;
        int clt = 0, cle = 0;
        for (int i = lane; i < c; i += 32) { const unsigned k = L[i].x; clt += (k < t); cle += (k <= t); }
        clt = __reduce_add_sync(0xffffffffu, clt);
        cle = __reduce_add_sync(0xffffffffu, cle);
        if (cle > RAST_K) {
            // ties on the depth at the cut: keep the (K - clt) lowest face ids among key == t
            const int need = RAST_K - clt;
            unsigned flo = 0u, fhi = 0xffffu;
            while (flo < fhi) {
                const unsigned mid = flo + ((fhi - flo) >> 1);
                int n = 0;
                for (int i = lane; i < c; i += 32) { const uint4 e = L[i]; n += (e.x == t && e.z <= mid) ? 1 : 0; }
                n = __reduce_add_sync(0xffffffffu, n);
                if (n >= need) fhi = mid; else flo = mid + 1;
            }
            tf = flo;
        }
    }
    float prod = 1.f;
    for (int i = lane; i < c; i += 32) {
        const uint4 e = L[i];
        if (e.x < t || (e.x == t && e.z <= tf)) prod *= __uint_as_float(e.y);
    }
    tkey = t; tfid = tf;
    return warp_prod(prod);
}

__global__ void __launch_bounds__(RT_THREADS, RT_CTAS_PER_SM)
raster_tile_forward_kernel(ModelDev m, Workspace w, TileScratch ts, int frame0, int n_frames, Weights wt, float* alpha_out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    RtSmem& sm = *reinterpret_cast<RtSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int S = w.S;
    const float inv_s = 1.f / (float)S;
    const int T = w.tiles_x * w.tiles_y;
    const int R = T * REGIONS_PER_TILE;
    const unsigned n_items = (unsigned)n_frames * (unsigned)T;
    const unsigned cap = (unsigned)ts.list_cap;
    uint4* list = ts.list + (size_t)blockIdx.x * ts.list_stride;
    unsigned* plane = sm.plane[wid];
    if (lane == 0) { mbar_init(&sm.bar[wid][0], 1); mbar_init(&sm.bar[wid][1], 1); mbar_fence_init(); }
    unsigned phase = 0;                  // bit b: parity the warp waits for next on its stage buffer b
    unsigned long long n_capped = 0, n_big = 0;

    for (;;) {
        __syncthreads();                 // the previous item's shared state is no longer read
        if (tid == 0) sm.item = (int)atomicAdd(ts.item_next, 1u);
        __syncthreads();
        const unsigned item = (unsigned)sm.item;
        if (item >= n_items) break;
        const int rank = (int)(item / (unsigned)n_frames), f = (int)(item % (unsigned)n_frames), fr = frame0 + f;
        const int tile = (int)w.tile_order[(size_t)fr * T + rank];
        const unsigned* toff = w.tile_off + (size_t)fr * (T + 1);
        const unsigned off = min(toff[tile], (unsigned)w.pool_cap);
        const int len = (int)(min(toff[tile + 1], (unsigned)w.pool_cap) - off);
        const int x0 = (tile % w.tiles_x) * TILE_W, y0 = (tile / w.tiles_x) * TILE_H;
        const size_t slot0 = ((size_t)fr * R + (size_t)tile * REGIONS_PER_TILE) * REGION_H;
        if (len == 0) {
            // no face reaches the tile: alpha = 0, |alpha - T| = T; pix is never read here
            if (tid < REGIONS_PER_TILE * REGION_H) w.region_l1[slot0 + tid] = w.region_tsum[slot0 + tid];
            if (alpha_out) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int x = x0 + lane, y = y0 + wid + RT_WARPS * k;
                    if (x < S && y < S) alpha_out[((size_t)f * S + y) * S + x] = 0.f;
                }
            }
            continue;
        }
        const int per = (len + RT_WARPS - 1) / RT_WARPS;
        const int lo = min(wid * per, len), hi = min(lo + per, len);
        const uint4* pool = w.tile_pool + (size_t)fr * w.pool_cap + off;
        const float4* recs = w.tile_rec + ((size_t)fr * w.pool_cap + off) * 4;

        for (unsigned round = 0;; ++round) {
            // ---- P0: candidates per (warp, pixel) = integral of the rectangles' corner grid
            for (int i = lane; i < RT_PLANE; i += 32) plane[i] = 0u;
            __syncwarp();
            for (int e = lo + lane; e < hi; e += 32) {
                const unsigned rect = pool[e].z;
                const int c0 = (int)(rect & 0xffu), c1 = (int)((rect >> 8) & 0xffu), r0 = (int)((rect >> 16) & 0xffu), r1 = (int)(rect >> 24);
                atomicAdd(&plane[r0 * RT_PITCH + c0], 1u);
                atomicAdd(&plane[r0 * RT_PITCH + c1 + 1], 0xffffffffu);
                atomicAdd(&plane[(r1 + 1) * RT_PITCH + c0], 0xffffffffu);
                atomicAdd(&plane[(r1 + 1) * RT_PITCH + c1 + 1], 1u);
            }
            __syncwarp();
            {
                unsigned run = 0u;
#pragma unroll 8
                for (int r = 0; r < TILE_H; ++r) { run += plane[r * RT_PITCH + lane]; plane[r * RT_PITCH + lane] = run; }
            }
            __syncwarp();
            {
                unsigned run = 0u;
#pragma unroll 8
                for (int c = 0; c < TILE_W; ++c) { run += plane[lane * RT_PITCH + c]; plane[lane * RT_PITCH + c] = run; }
            }
            __syncthreads();

            // ---- P0b: classify the pixels (thread: column = lane, rows wid + 8k), lay out the lists
            unsigned cnt4[4];
            unsigned mysum = 0u;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int idx = (wid + RT_WARPS * k) * RT_PITCH + lane;
                unsigned c = 0u;
#pragma unroll
                for (int q = 0; q < RT_WARPS; ++q) c += sm.plane[q][idx];
                cnt4[k] = c;
                if (c > (unsigned)RAST_K) mysum += c;
            }
            unsigned incl = mysum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const unsigned y = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += y; }
            if (lane == 31) sm.warp_tot[wid] = incl;
            if (tid == 0) sm.n_active = 0u;
            __syncthreads();
            unsigned base = 0u, total = 0u;
#pragma unroll
            for (int q = 0; q < RT_WARPS; ++q) { const unsigned v = sm.warp_tot[q]; if (q < wid) base += v; total += v; }
            unsigned run = base + incl - mysum;
            const unsigned win_lo = round * cap;
            unsigned actmask = 0u;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int row = wid + RT_WARPS * k;
                const int idx = row * RT_PITCH + lane, px = row * TILE_W + lane;
                const unsigned c = cnt4[k];
                if (c > (unsigned)RAST_K) {
                    const unsigned o = run;
                    run += c;
                    const bool act = (o >= win_lo) && (o - win_lo < cap);
                    unsigned cur = RT_LISTED | (o - win_lo);
#pragma unroll
                    for (int q = 0; q < RT_WARPS; ++q) {
                        const unsigned cw = sm.plane[q][idx];
                        sm.plane[q][idx] = act ? cur : RT_SKIP;
                        cur += cw;
                    }
                    if (act) {
                        const unsigned a = atomicAdd(&sm.n_active, 1u);
                        sm.active[a] = (unsigned short)px;
                        sm.list_off[px] = o - win_lo;
                        sm.list_cnt[px] = (unsigned short)c;
                        actmask |= 1u << k;
                    }
                } else {
                    const unsigned v = (round == 0u) ? 0x3f800000u : RT_SKIP;
#pragma unroll
                    for (int q = 0; q < RT_WARPS; ++q) sm.plane[q][idx] = v;
                }
            }
            __syncthreads();

            // ---- P1: sweep this warp's faces (TMA double buffer of RT_BLK prepared faces)
            {
                const int nblk = (hi - lo + RT_BLK - 1) / RT_BLK;
                if (nblk > 0 && lane == 0) {
                    // the stage doubles as P2's scratch (generic-proxy writes): order them before the bulk copies
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    const unsigned bytes = (unsigned)min(RT_BLK, hi - lo) * 64u;
                    mbar_expect_tx(&sm.bar[wid][0], bytes);
                    tma_load_1d(sm.stage[wid][0], recs + (size_t)lo * 4, bytes, &sm.bar[wid][0]);
                }
                for (int b = 0; b < nblk; ++b) {
                    const int e0 = lo + b * RT_BLK;
                    if (b + 1 < nblk && lane == 0) {
                        const unsigned bytes = (unsigned)min(RT_BLK, hi - (e0 + RT_BLK)) * 64u;
                        mbar_expect_tx(&sm.bar[wid][(b + 1) & 1], bytes);
                        tma_load_1d(sm.stage[wid][(b + 1) & 1], recs + (size_t)(e0 + RT_BLK) * 4, bytes, &sm.bar[wid][(b + 1) & 1]);
                    }
                    mbar_wait(&sm.bar[wid][b & 1], (phase >> (b & 1)) & 1u);
                    phase ^= 1u << (b & 1);
                    const int nrec = min(RT_BLK, hi - e0);
                    const float4* st = sm.stage[wid][b & 1];
                    for (int j = 0; j < nrec; ++j) {
                        const float4 q0 = st[j * 4 + 0], q1 = st[j * 4 + 1], q2 = st[j * 4 + 2], q3 = st[j * 4 + 3];
                        FaceSetup fs;
                        fs.x0 = q0.x; fs.y0 = q0.y; fs.x1 = q0.z; fs.y1 = q0.w; fs.x2 = q1.x; fs.y2 = q1.y;
                        fs.z0 = q1.z; fs.z1 = q1.w; fs.z2 = q2.x; fs.rden = q2.y;
                        fs.rl01 = q2.z; fs.rl02 = q2.w; fs.rl12 = q3.x;
                        fs.e01x = fsub(fs.x1, fs.x0); fs.e01y = fsub(fs.y1, fs.y0);
                        fs.e02x = fsub(fs.x2, fs.x0); fs.e02y = fsub(fs.y2, fs.y0);
                        fs.e12x = fsub(fs.x2, fs.x1); fs.e12y = fsub(fs.y2, fs.y1);
                        const unsigned fid = __float_as_uint(q3.y), rect = __float_as_uint(q3.z);
                        const int c0 = (int)(rect & 0xffu), c1 = (int)((rect >> 8) & 0xffu), r0 = (int)((rect >> 16) & 0xffu), r1 = (int)(rect >> 24);
                        const int wd = c1 - c0 + 1, npx = wd * (r1 - r0 + 1);
                        const float inv_w = 1.f / (float)wd;
                        for (int i = lane; i < npx; i += 32) {
                            const int rr = (int)(((float)i + 0.5f) * inv_w);
                            const int lx = c0 + (i - rr * wd), ly = r0 + rr;
                            const int idx = ly * RT_PITCH + lx;
                            const unsigned v = plane[idx];
                            if (v == RT_SKIP) continue;
                            float sd, pz, mv = 1.f;
                            const bool ok = frag_setup_forward(fs, pix_to_ndc(x0 + lx, inv_s), pix_to_ndc(y0 + ly, inv_s), sd, pz);
                            if (ok) { float pp; frag_prob(sd, pp, mv); }
                            if (v & RT_LISTED) {
                                plane[idx] = v + 1u;
                                list[v & 0x7fffffffu] = make_uint4(ok ? __float_as_uint(pz + 0.f) : 0xffffffffu, __float_as_uint(mv), fid, 0u);
                            } else if (ok) {
                                plane[idx] = __float_as_uint(__uint_as_float(v) * mv);
                            }
                        }
                        __syncwarp();        // the next face's lanes may touch the same pixels
                    }
                }
            }
            __syncthreads();

            // ---- P2: listed pixels of this pass, one warp each
            {
                const unsigned na = sm.n_active;
                unsigned* scratch = reinterpret_cast<unsigned*>(sm.stage[wid][0]);     // the stage is idle here
                for (unsigned a = (unsigned)wid; a < na; a += RT_WARPS) {
                    const int px = (int)sm.active[a];
                    const int c = (int)sm.list_cnt[px];
                    unsigned tk, tf;
                    bool capped;
                    const float P = rt_select(list + sm.list_off[px], c, lane, scratch, tk, tf, capped);
                    __syncwarp();
                    if (lane == 0) {
                        sm.plane[0][(px / TILE_W) * RT_PITCH + (px % TILE_W)] = __float_as_uint(P);
                        sm.list_off[px] = tk;
                        sm.list_cnt[px] = (unsigned short)tf;
                        if (capped) ++n_capped;
                        if (c > RT_SELCAP) ++n_big;
                    }
                }
            }
            __syncthreads();

            // ---- P3: finish the pixels of this pass
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int row = wid + RT_WARPS * k;
                const int idx = row * RT_PITCH + lane, px = row * TILE_W + lane;
                const bool listed = cnt4[k] > (unsigned)RAST_K;
                if (listed ? !((actmask >> k) & 1u) : (round != 0u)) continue;
                float P = 1.f;
                unsigned tk = 0xffffffffu, tf = 0xffffu;
                if (listed) {
                    P = __uint_as_float(sm.plane[0][idx]); tk = sm.list_off[px]; tf = sm.list_cnt[px];
                } else if (cnt4[k] > 0u) {
#pragma unroll
                    for (int q = 0; q < RT_WARPS; ++q) P *= __uint_as_float(sm.plane[q][idx]);
                }
                const int x = x0 + lane, y = y0 + row;
                float l1 = 0.f;
                if (x < S && y < S) {
                    const size_t pi = ((size_t)fr * S + y) * S + x;
                    const float alpha = 1.f - P;
                    const float d = alpha - (float)w.sil[pi];
                    l1 = fabsf(d);
                    float coef = 0.f;
                    if (P < 1.f && P >= P_SKIP && d != 0.f) {        // P < 1 <=> the pixel has fragments
                        const float ga = wt.sil * w.inv_window[fr] * inv_s * inv_s * (d > 0.f ? 1.f : -1.f);
                        coef = ga * P * (1.f / RAST_SIGMA);
                    }
                    w.pix[pi] = make_uint2(__float_as_uint(coef), tk);
                    if (tk != 0xffffffffu) w.pix_tfid[pi] = (unsigned short)tf;
                    if (alpha_out) alpha_out[((size_t)f * S + y) * S + x] = alpha;
                }
                sm.l1[px] = l1;
            }
            if (total <= (round + 1u) * cap) break;      // every listed pixel started inside a window already done
            __syncthreads();                             // planes are rebuilt by the next pass
        }
        __syncthreads();
        // per region row (8 pixels), fixed order
        if (tid < REGIONS_PER_TILE * REGION_H) {
            const int sub = tid / REGION_H, row = tid % REGION_H;
            const int lx0 = (sub % (TILE_W / REGION_W)) * REGION_W, ly0 = (sub / (TILE_W / REGION_W)) * REGION_H;
            const float* p = sm.l1 + (ly0 + row) * TILE_W + lx0;
            w.region_l1[slot0 + tid] = ((p[0] + p[1]) + (p[2] + p[3])) + ((p[4] + p[5]) + (p[6] + p[7]));
        }
    }
    if (lane == 0 && (n_capped | n_big)) {
        atomicAdd(w.counters + 0, n_capped);
        atomicAdd(w.counters + 1, n_big);
    }
}

size_t raster_tile_smem_bytes() { return sizeof(RtSmem); }

void launch_raster_tile_forward(const ModelDev& m, const Workspace& w, const TileScratch& ts, int frame0, int n, Weights wt,
                                float* alpha_out, int n_ctas, cudaStream_t st) {
    cudaMemsetAsync(ts.item_next, 0, sizeof(unsigned), st);
    const long long items = (long long)n * w.tiles_x * w.tiles_y;
    const int grid = (int)(items < n_ctas ? items : n_ctas);
    raster_tile_forward_kernel<<<grid, RT_THREADS, raster_tile_smem_bytes(), st>>>(m, w, ts, frame0, n, wt, alpha_out);
}

}  // namespace smf
