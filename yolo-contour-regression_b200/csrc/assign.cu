// Training path, assignment stage: in-box candidate rectangles (K0), polar targets + Polar-IoU +
// align metric per candidate (K1), per-GT top-k (K2), per-image multi-GT resolution / targets /
// normalisation (K3), polar targets (+ Polar-IoU loss and its gradient) of the positives (K4),
// and the dense API outputs of TaskAlignedAssigner.forward.
//
// Reference: utils/tal.py:1135-1390, :52-66, :214-248, :1445-1464 (file:line under
// /root/reference/ultralytics-main/ultralytics/).  No (B,G,A) float tensor is materialised.
#include "train_path.cuh"
#include <stdlib.h>

// Candidates per chunk = threads per block of K1.  One-warp blocks (nine per SM at 36 rays) need no barrier
// partner and no queue sharing and measure 0.6 % faster than two-warp blocks (five per SM, queues shared
// between the warps); -DK1_NT=64 selects the latter.
#ifndef K1_NT
#define K1_NT 32
#endif
#ifndef K1_NT_NARROW
#define K1_NT_NARROW 32   // 72 rays: 1.2 KB of lists per thread - one-warp blocks fit five per SM (two-warp blocks: two)
#endif
// threads per block of K1 = candidates per chunk, by ray count
static inline int k1_nt(int R) { return (R > 36) ? K1_NT_NARROW : K1_NT; }
template <int R> struct K1Nt { static constexpr int value = (R > 36) ? K1_NT_NARROW : K1_NT; };
#ifndef K1_SHARE_QUEUES
#define K1_SHARE_QUEUES 1
#endif
#ifndef YCR_STATS
#define YCR_STATS 0
#endif
#define K3_NT 512
#define K4_NT 32   // positives per GT are ~topk: one warp per block

// running totals since the last ycr_debug_stats(reset): candidates, pairs queued for neighbourhood
// settlement, pairs that needed the exact scan (measurement aid, three atomics per block iteration)
__device__ unsigned long long g_ycr_stats[4];

int debug_stats(unsigned long long* out_h, int reset) {
    YCR_CUDA_CHECK(cudaDeviceSynchronize());
    YCR_CUDA_CHECK(cudaMemcpyFromSymbol(out_h, g_ycr_stats, sizeof(unsigned long long) * 4));
    if (reset) {
        unsigned long long z[4] = {0, 0, 0, 0};
        YCR_CUDA_CHECK(cudaMemcpyToSymbol(g_ycr_stats, z, sizeof(z)));
    }
    return YCR_OK;
}

// ------------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------------
struct AnchorPos { int level, iy, ix, a_local; };

__device__ __forceinline__ AnchorPos anchor_pos(const GridDev& g, int a) {
    AnchorPos p;
    int l = 0;
#pragma unroll
    for (int k = 1; k < YCR_MAX_LEVELS; ++k)
        if (k < g.n_levels && a >= g.off[k]) l = k;
    p.level = l;
    p.a_local = a - g.off[l];
    p.iy = p.a_local / g.w[l];
    p.ix = p.a_local - p.iy * g.w[l];
    return p;
}

// candidate c of GT bg -> anchor (level-major, row-major inside each level's rectangle)
__device__ __forceinline__ AnchorPos cand_anchor(const GridDev& g, const int4* rect, int c) {
    AnchorPos p{0, 0, 0, 0};
    for (int l = 0; l < g.n_levels; ++l) {
        const int4 r = rect[l];
        const int n = r.z * r.w;
        if (c < n) {
            const int yy = c / r.z;
            p.level = l;
            p.iy = r.y + yy;
            p.ix = r.x + (c - yy * r.z);
            p.a_local = p.iy * g.w[l] + p.ix;
            return p;
        }
        c -= n;
    }
    return p;
}

// candidate index of anchor position p inside GT bg, or -1 when the anchor is not in the GT's box
__device__ __forceinline__ int cand_index(const GridDev& g, const int4* rect, const AnchorPos& p) {
    int base = 0;
    for (int l = 0; l < p.level; ++l) base += rect[l].z * rect[l].w;
    const int4 r = rect[p.level];
    const int dx = p.ix - r.x, dy = p.iy - r.y;
    if (dx < 0 || dy < 0 || dx >= r.z || dy >= r.w) return -1;
    return base + dy * r.z + dx;
}

__device__ __forceinline__ float anchor_coord(int i, float stride) { return ((float)i + 0.5f) * stride; }

// ------------------------------------------------------------------------------------------------
// K0: candidate rectangles.  The in-box predicate of select_candidates_in_gts (utils/tal.py:52-66),
// min(ax-x1, ay-y1, x2-ax, y2-ay) > 1e-9, is separable in x and y, so the candidate set of a GT on
// each level is exactly a rectangle of grid cells, found with the same fp32 comparisons.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void axis_range(float lo_edge, float hi_edge, float stride, int n, int& first, int& count) {
    const float eps = 1e-9f;
    int lo = (int)floorf(lo_edge / stride - 0.5f);
    lo = max(0, min(n, lo));
    while (lo > 0 && (anchor_coord(lo - 1, stride) - lo_edge) > eps) --lo;
    while (lo < n && !((anchor_coord(lo, stride) - lo_edge) > eps)) ++lo;
    int hi = (int)ceilf(hi_edge / stride - 0.5f);
    hi = max(-1, min(n - 1, hi));
    while (hi < n - 1 && (hi_edge - anchor_coord(hi + 1, stride)) > eps) ++hi;
    while (hi >= 0 && !((hi_edge - anchor_coord(hi, stride)) > eps)) --hi;
    first = lo;
    count = max(0, hi - lo + 1);
}

// Candidate rectangles, validity and candidate count of every GT (one thread per GT, any number of blocks).
__global__ void __launch_bounds__(128) k_gt_rects(GridDev grid, ycr_gt_t gt, AssignWs ws) {
    pdl_enter();
    const int BG = gt.B * gt.G;
    const int bg = blockIdx.x * 128 + threadIdx.x;
    if (bg >= BG) return;
    const float* bx = gt.boxes + (int64_t)bg * gt.boxes_stride;
    const float x1 = bx[0], y1 = bx[1], x2 = bx[2], y2 = bx[3];
    const bool valid = gt.mask_gt ? (gt.mask_gt[(int64_t)bg * gt.mask_stride] != 0.f) : ((x1 + y1 + x2 + y2) > 0.f);
    int n = 0;
    for (int l = 0; l < grid.n_levels; ++l) {
        int fx = 0, cx = 0, fy = 0, cy = 0;
        if (valid) {
            axis_range(x1, x2, grid.stride[l], grid.w[l], fx, cx);
            axis_range(y1, y2, grid.stride[l], grid.h[l], fy, cy);
            if (cx == 0 || cy == 0) cx = cy = 0;
        }
        ws.rect[bg * YCR_MAX_LEVELS + l] = make_int4(fx, fy, cx, cy);
        n += cx * cy;
    }
    ws.valid[bg] = valid ? 1 : 0;
    ws.ncand[bg] = n;
}

// One block; GT i of each round of 1024 is handled by thread i, the running candidate / chunk / partial-chunk
// offsets are block-wide exclusive scans (warp shuffles + shared arrays) carried from round to round.  The
// last phase lays out the order in which K1 draws the chunks: the full chunks of all GTs first, then the
// partial chunk of every GT that has one - the cheapest work last keeps the tail of the persistent kernel short.
__global__ void __launch_bounds__(1024) k_gt_setup(ycr_gt_t gt, AssignWs ws, int chunk) {
    pdl_enter();
    __shared__ int s_wc[32], s_wk[32], s_wp[32];
    __shared__ int s_carry[3];
    __shared__ int s_n[1024], s_first[1024], s_po[1024];
    const int BG = gt.B * gt.G;
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    if (t == 0) { s_carry[0] = 0; s_carry[1] = 0; s_carry[2] = 0; }
    __syncthreads();
    for (int base = 0; base < BG; base += 1024) {
        const int bg = base + t;
        int n = 0, nk = 0, np = 0;
        if (bg < BG) {
            n = ws.ncand[bg];   // k_gt_rects
            nk = (n + chunk - 1) / chunk;
            np = (n % chunk) ? 1 : 0;
        }
        // inclusive scans inside the warp, then across the 32 warp totals
        int ic = n, ik = nk, ip = np;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int vc = __shfl_up_sync(0xffffffffu, ic, o), vk = __shfl_up_sync(0xffffffffu, ik, o);
            const int vp = __shfl_up_sync(0xffffffffu, ip, o);
            if (lane >= o) { ic += vc; ik += vk; ip += vp; }
        }
        if (lane == 31) { s_wc[wid] = ic; s_wk[wid] = ik; s_wp[wid] = ip; }
        __syncthreads();
        if (wid == 0) {
            int wc = s_wc[lane], wk = s_wk[lane], wp = s_wp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int vc = __shfl_up_sync(0xffffffffu, wc, o), vk = __shfl_up_sync(0xffffffffu, wk, o);
                const int vp = __shfl_up_sync(0xffffffffu, wp, o);
                if (lane >= o) { wc += vc; wk += vk; wp += vp; }
            }
            s_wc[lane] = wc;
            s_wk[lane] = wk;
            s_wp[lane] = wp;
        }
        __syncthreads();
        const int run_c = s_carry[0] + (wid ? s_wc[wid - 1] : 0) + ic - n;
        const int run_k = s_carry[1] + (wid ? s_wk[wid - 1] : 0) + ik - nk;
        const int run_p = s_carry[2] + (wid ? s_wp[wid - 1] : 0) + ip - np;
        if (bg < BG) {
            ws.cand_off[bg] = run_c;
            ws.chunk_off[bg] = run_k;
            ws.part_off[bg] = run_p;
        }
        s_n[t] = nk - np;                 // full chunks of this GT
        s_first[t] = run_k;
        s_po[t] = run_k - run_p;          // full chunks of all GTs before it = its first hand-out slot
        __syncthreads();
        if (t == 1023) { s_carry[0] = run_c + n; s_carry[1] = run_k + nk; s_carry[2] = run_p + np; }
        // hand-out entries of the full chunks of this round's GTs: one thread per entry, the GT found by
        // binary search in the (non-decreasing) first slots - the owner of slot u is the last GT whose first
        // slot is <= u
        {
            const int nround = min(1024, BG - base);
            const int u_begin = s_po[0], u_end = s_po[nround - 1] + s_n[nround - 1];
            for (int u = u_begin + t; u < u_end; u += 1024) {
                int lo = 0, hi = nround - 1;
                while (lo < hi) {
                    const int mid = (lo + hi + 1) >> 1;
                    if (s_po[mid] <= u) lo = mid; else hi = mid - 1;
                }
                if (u < ws.chunks_cap) { ws.chunk_bg[u] = base + lo; ws.chunk_work[u] = s_first[lo] + (u - s_po[lo]); }
            }
        }
        __syncthreads();
    }
    const int M = s_carry[0], T = s_carry[1], P = s_carry[2];
    for (int bg = t; bg < BG; bg += 1024) {   // the partial chunks go behind all full ones
        const int n = ws.ncand[bg];
        if (n % chunk) {
            const int u = T - P + ws.part_off[bg];
            if (u < ws.chunks_cap) { ws.chunk_bg[u] = bg; ws.chunk_work[u] = ws.chunk_off[bg] + n / chunk; }
        }
    }
    if (t == 0) {
        ws.cand_off[BG] = M;
        ws.chunk_off[BG] = T;
        ws.totals[0] = M;
        ws.totals[1] = T;
        ws.totals[2] = 0;  // K1's work counter
        ws.totals[3] = 0;  // K3's finished-block counter
        ws.err[0] = ((int64_t)M > ws.cand_cap) ? 1 : 0;
    }
}

// ------------------------------------------------------------------------------------------------
// K1: one thread per in-box candidate (b,g,a): polar targets, Polar-IoU (MaskIOU utils/tal.py:1445),
// align metric score^alpha * ov^beta (utils/tal.py:1281).  Persistent blocks walk the chunk list.
// ------------------------------------------------------------------------------------------------
template <int R, int NT>
__device__ __forceinline__ void init_raydir(PolarSmem<R, NT>& sm, int tid) {
    for (int i = tid; i < R; i += NT) {
        const double ang = (double)(i * (360 / R)) * (3.14159265358979323846 / 180.0);
        sm.raydir[i] = make_float2((float)cos(ang), (float)sin(ang));
    }
}

__device__ __forceinline__ float align_of(float score, float ov, float alpha, float beta) {
    const float s = (alpha == 0.5f) ? sqrtf(score) : powf(score, alpha);
    float o;
    if (beta == 4.0f) { const float o2 = ov * ov; o = o2 * o2; }
    else o = powf(ov, beta);
    return s * o;
}

template <int R, int NT>
__global__ void __launch_bounds__(NT) k_cand_overlaps(const __grid_constant__ AssignArgs a, const __grid_constant__ AssignWs ws) {
    pdl_enter();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PolarSmem<R, NT>& sm = *reinterpret_cast<PolarSmem<R, NT>*>(smem_raw);
    // chunk descriptor of the current iteration: [0] first chunk of the GT, [1] its candidate count,
    // [2] chunk index, [3] GT index, [4..] the GT's candidate rectangles (int4 per level)
    __shared__ __align__(16) int s_desc[4 + 4 * YCR_MAX_LEVELS];
    __shared__ int s_ctl[6];   // queue sharing between the two warps (polar_settle_queue_shared)
    static_assert(4 + 4 * YCR_MAX_LEVELS <= 32, "descriptor is fetched by one warp, one word per lane");
    constexpr int CW = (2 * YCR_C + 31) / 32;   // contour words per lane of warp 0
    const int tid = threadIdx.x;
    const int T = ws.totals[1];
    if (ws.err[0]) return;
    init_raydir<R, NT>(sm, tid);

    // Blocks draw chunks from a global counter (k_gt_setup zeroes it): chunk costs vary a lot with the GT's
    // shape, and a static split leaves blocks idle at the end.  Warp 0 runs one chunk ahead: it draws the
    // next chunk before the sweep and fetches that chunk's descriptor and contour into registers after it,
    // so the loads are in flight during the settlement and land in shared memory at the next barrier.
    int n_work = T, n_bg = -1, n_desc = 0;
    float n_contour[CW];
    auto fetch_next = [&](int drawn) {          // warp 0, all lanes; `drawn` valid in lane 0
        const int unit = __shfl_sync(0xffffffffu, drawn, 0);
        n_work = T;
        n_bg = -1;
        if (unit < T) {
            int bgl = 0, wl = 0;
            if (tid == 0) { bgl = ws.chunk_bg[unit]; wl = ws.chunk_work[unit]; }
            n_bg = __shfl_sync(0xffffffffu, bgl, 0);
            n_work = __shfl_sync(0xffffffffu, wl, 0);
            const float* cp = a.gt.coor + (int64_t)n_bg * a.gt.coor_stride;
#pragma unroll
            for (int k = 0; k < CW; ++k)
                if (k * 32 + tid < 2 * YCR_C) n_contour[k] = cp[k * 32 + tid];
            if (tid == 0) n_desc = ws.chunk_off[n_bg];
            else if (tid == 1) n_desc = ws.ncand[n_bg];
            else if (tid == 2) n_desc = n_work;
            else if (tid == 3) n_desc = n_bg;
            else if (tid < 4 + 4 * YCR_MAX_LEVELS)
                n_desc = reinterpret_cast<const int*>(ws.rect)[n_bg * 4 * YCR_MAX_LEVELS + tid - 4];
        } else if (tid == 2) {
            n_desc = T;
        }
    };
    if (tid < 32) {
        int drawn = 0;
        if (tid == 0) drawn = atomicAdd(&ws.totals[2], 1);
        fetch_next(drawn);
    }
    for (;;) {
        __syncthreads();   // both warps are done with the previous chunk's contour and descriptor
        if (tid < 32) {
            if (n_bg >= 0) {
                float* dst = reinterpret_cast<float*>(sm.contour);
#pragma unroll
                for (int k = 0; k < CW; ++k)
                    if (k * 32 + tid < 2 * YCR_C) dst[k * 32 + tid] = n_contour[k];
            }
            if (tid < 4 + 4 * YCR_MAX_LEVELS) s_desc[tid] = n_desc;
            if (tid < 6) s_ctl[tid] = (tid < 2) ? -1 : 0;
        }
        __syncthreads();
        const int work = s_desc[2];
        if (work >= T) break;
        const int bg = s_desc[3];
        int drawn = 0;
        if (tid == 0) drawn = atomicAdd(&ws.totals[2], 1);   // consumed after the sweep
        const int c = (work - s_desc[0]) * NT + tid;
        const int ncand = s_desc[1];
        const bool active = c < ncand;
        AnchorPos ap{0, 0, 0, 0};
        float ax = 0.f, ay = 0.f;
        // the candidate's predicted rays and class score are requested now and used after the settlement
        uint32_t pr[R];      // raw words: converted where they are used, so that no load waits for its data here
        uint32_t score_raw = 0;
        float score = 0.f;
        if (active) {
            ap = cand_anchor(a.grid, reinterpret_cast<const int4*>(&s_desc[4]), c);
            ax = anchor_coord(ap.ix, a.grid.stride[ap.level]);
            ay = anchor_coord(ap.iy, a.grid.stride[ap.level]);
            const int b = bg / a.gt.G;
            const int l = ap.level;
            const int64_t r0 = (int64_t)b * a.pred.rays_sb[l] + (int64_t)ap.a_local * a.pred.rays_sa[l];
            const int64_t sc = a.pred.rays_sc[l];
            const int label = (int)a.gt.labels[(int64_t)bg * a.gt.labels_stride];
            const int64_t si = (int64_t)b * a.pred.cls_sb[l] + (int64_t)ap.a_local * a.pred.cls_sa[l] + (int64_t)label * a.pred.cls_sc[l];
            if (a.dtype == YCR_F32) {   // (one branch around the whole batch of loads, not one per load)
                const uint32_t* rp = reinterpret_cast<const uint32_t*>(a.pred.rays[l]) + r0;
#pragma unroll
                for (int i = 0; i < R; ++i) pr[i] = rp[i * sc];
                score_raw = reinterpret_cast<const uint32_t*>(a.pred.cls[l])[si];
            } else {
                const unsigned short* rp = reinterpret_cast<const unsigned short*>(a.pred.rays[l]) + r0;
#pragma unroll
                for (int i = 0; i < R; ++i) pr[i] = rp[i * sc];
                score_raw = reinterpret_cast<const unsigned short*>(a.pred.cls[l])[si];
            }
            polar_sweep<R, NT>(sm, a.pc, tid, ax, ay);
        }
        if (tid < 32) fetch_next(drawn);
        const int nq = polar_settle_own<R, NT>(sm, a.pc, tid, active, ax, ay);
        int nscan;
        if constexpr (NT == 64 && K1_SHARE_QUEUES) nscan = polar_settle_queue_shared<R, NT>(sm, a.pc, tid, nq, s_ctl);
        else nscan = polar_settle_queue<R, NT>(sm, a.pc, tid, nq);
#if YCR_STATS   // measurement builds only (YCR_NVCC_FLAGS=-DYCR_STATS=1): three global atomics per warp and chunk
        if ((tid & 31) == 0) {
            atomicAdd(&g_ycr_stats[0], (unsigned long long)max(0, min(32, ncand - (work - s_desc[0]) * NT - (tid & ~31))));
            atomicAdd(&g_ycr_stats[1], (unsigned long long)nq);
            atomicAdd(&g_ycr_stats[2], (unsigned long long)nscan);
        }
#else
        (void)nscan;
#endif
        if (active) {
            const float rs = a.pred.ray_scale[ap.level];
            float smin = 0.f, smax = 0.f;
#pragma unroll
            for (int i = 0; i < R; ++i) {
                const float p = ycr_round_to(ycr_from_raw(pr[i], a.dtype) * rs, a.dtype);   // the reference multiplies in the input type
                const float t = sm.tv(i, tid);
                smin += fmaxf(fminf(p, t), YCR_FLOOR);
                smax += fmaxf(p, t);
            }
            const float ov = smin / smax;
            // `pred_scores.detach().sigmoid()` stays in the input type (utils/loss.py:861)
            score = ycr_from_raw(score_raw, a.dtype);
            if (a.pred.cls_is_logit) score = ycr_round_to(1.f / (1.f + expf(-score)), a.dtype);
            if (ws.cand_t) {
                float* tp = ws.cand_t + (int64_t)work * R * NT + tid;
#pragma unroll 4
                for (int i = 0; i < R; ++i) tp[i * NT] = sm.tv(i, tid);
            }
            const int64_t m = (int64_t)ws.cand_off[bg] + c;
            ws.cand_ov[m] = ov;
            ws.cand_align[m] = align_of(score, ov, a.cfg.alpha, a.cfg.beta);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K2: per-GT top-k over its candidates, one warp per GT (select_topk_candidates utils/tal.py:1304).
// Order: metric descending, anchor index ascending on ties (candidates are enumerated in anchor
// order).  If fewer than topk candidates have a positive metric, torch.topk over the full anchor
// axis pads with zero-metric anchors; with lowest-index tie-breaking an in-box zero-metric candidate
// is picked iff fewer than (topk - n_pos) zero-metric anchors precede it.
// ------------------------------------------------------------------------------------------------
#define K2_CACHE 2048   // align metrics of one GT kept in shared memory (per warp); larger GTs re-read L2
__global__ void __launch_bounds__(128) k_topk_per_gt(const __grid_constant__ AssignArgs a, const __grid_constant__ AssignWs ws) {
    pdl_enter();
    __shared__ float s_al[4][K2_CACHE];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bg = blockIdx.x * 4 + warp;
    const int BG = a.gt.B * a.gt.G;
    if (bg >= BG) return;
    const int topk = a.cfg.topk;
    int* sel = ws.sel + (int64_t)bg * topk;
    if (ws.err[0]) {   // candidate capacity exceeded: K1 did not run, the per-candidate arrays are too short to index
        for (int k = lane; k < topk; k += 32) sel[k] = -1;
        return;
    }
    const int n = ws.valid[bg] ? ws.ncand[bg] : 0;
    const float* al = ws.cand_align + ws.cand_off[bg];
    const int4* rect = ws.rect + bg * YCR_MAX_LEVELS;
    int n_sel = 0;
    int my_pick = -1;   // lane k keeps the k-th pick (topk <= 32), converted to an anchor index at the end
    if (n <= K2_CACHE) {
        // cached form: a picked entry is overwritten with -1, so every pass is a plain arg-max
        float* s = s_al[warp];
        for (int c = lane; c < n; c += 32) s[c] = al[c];
        __syncwarp();
        for (int k = 0; k < topk && n > 0; ++k) {
            float bv = 0.f;
            int bc = 0x7fffffff;
#pragma unroll 4
            for (int c = lane; c < n; c += 32) {
                const float v = s[c];
                if (v > bv) { bv = v; bc = c; }  // ascending c: first maximum kept
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int oc = __shfl_xor_sync(0xffffffffu, bc, o);
                if (ov > bv || (ov == bv && oc < bc)) { bv = ov; bc = oc; }
            }
            if (bc == 0x7fffffff) break;
            if (lane == 0) s[bc] = -1.f;
            if (topk <= 32) { if (lane == n_sel) my_pick = bc; }
            else if (lane == 0) { const AnchorPos p = cand_anchor(a.grid, rect, bc); sel[n_sel] = a.grid.off[p.level] + p.a_local; }
            ++n_sel;
            __syncwarp();
        }
    } else {
        float last_v = __int_as_float(0x7f800000);  // +inf
        int last_c = -1;
        for (int k = 0; k < topk && n > 0; ++k) {
            float bv = 0.f;
            int bc = 0x7fffffff;
            for (int c = lane; c < n; c += 32) {
                const float v = al[c];
                const bool elig = (v > 0.f) && (v < last_v || (v == last_v && c > last_c));
                if (elig && (v > bv)) { bv = v; bc = c; }  // ascending c: first maximum kept
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int oc = __shfl_xor_sync(0xffffffffu, bc, o);
                if (ov > bv || (ov == bv && oc < bc)) { bv = ov; bc = oc; }
            }
            if (bc == 0x7fffffff) break;
            if (topk <= 32) { if (lane == n_sel) my_pick = bc; }
            else if (lane == 0) { const AnchorPos p = cand_anchor(a.grid, rect, bc); sel[n_sel] = a.grid.off[p.level] + p.a_local; }
            ++n_sel;
            last_v = bv;
            last_c = bc;
        }
    }
    if (topk <= 32 && lane < n_sel) {
        const AnchorPos p = cand_anchor(a.grid, rect, my_pick);
        sel[lane] = a.grid.off[p.level] + p.a_local;
    }
    if (n_sel < topk && n > 0) {
        // zero-metric in-box candidates as topk fillers (rare)
        const int want = topk - n_sel;
        for (int c0 = 0; c0 < n && n_sel < topk; c0 += 32) {
            const int c = c0 + lane;
            const bool zero = (c < n) && (al[c] == 0.f);
            unsigned ball = __ballot_sync(0xffffffffu, zero);
            while (ball && n_sel < topk) {
                const int src = __ffs(ball) - 1;
                ball &= ball - 1;
                const int cz = c0 + src;
                int cnt = 0;  // positive-metric candidates before cz
                for (int q = lane; q < cz; q += 32) cnt += (al[q] > 0.f) ? 1 : 0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
                const AnchorPos p = cand_anchor(a.grid, rect, cz);
                const int anchor = a.grid.off[p.level] + p.a_local;
                if (anchor - cnt < want) {
                    if (lane == 0) sel[n_sel] = anchor;
                    ++n_sel;
                }
            }
        }
    }
    for (int k = n_sel + lane; k < topk; k += 32) sel[k] = -1;
}

// ------------------------------------------------------------------------------------------------
// K3: one block per image.  Counts how many GTs picked each anchor, resolves anchors picked by
// several GTs to the GT with the highest overlap over ALL in-box GTs (select_highest_overlaps
// utils/tal.py:214-248), orders the positives (g,a)-lexicographically (utils/tal.py:1175), and
// computes the normalised target score (utils/tal.py:1197-1202).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(K3_NT) k_resolve_image(const __grid_constant__ AssignArgs a, const __grid_constant__ AssignWs ws,
                                                         int* n_pos_d) {
    pdl_enter();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int A = a.grid.off[YCR_MAX_LEVELS];
    const int G = a.gt.G, topk = a.cfg.topk, pos_cap = ws.pos_cap;
    int4* s_rect = reinterpret_cast<int4*>(smem_raw);               // [G][levels] candidate rectangles of the image's GTs
    uint32_t* s_a = reinterpret_cast<uint32_t*>(s_rect + G * YCR_MAX_LEVELS);  // [A]
    int* s_pa = reinterpret_cast<int*>(s_a + A);                    // [pos_cap]
    int* s_pg = s_pa + pos_cap;                                     // [pos_cap]
    float* s_al = reinterpret_cast<float*>(s_pg + pos_cap);         // [pos_cap]
    int* s_conf = reinterpret_cast<int*>(s_al + pos_cap);           // [pos_cap] anchors picked by several GTs
    uint32_t* s_gal = reinterpret_cast<uint32_t*>(s_conf + pos_cap);  // [G]
    uint32_t* s_gov = s_gal + G;                                    // [G]
    int* s_gcnt = reinterpret_cast<int*>(s_gov + G);                // [G]
    int* s_gstart = s_gcnt + G;                                     // [G]
    int* s_goff = s_gstart + G;                                     // [G] first candidate of the GT, -1 when invalid
    __shared__ int s_n, s_nc, s_last;
    __shared__ float s_red[32];
    const int b = blockIdx.x, tid = threadIdx.x;
    for (int i = tid; i < A; i += K3_NT) s_a[i] = 0;
    for (int i = tid; i < G; i += K3_NT) {
        s_gal[i] = 0; s_gov[i] = 0; s_gcnt[i] = 0;
        s_goff[i] = ws.valid[b * G + i] ? ws.cand_off[b * G + i] : -1;
    }
    for (int i = tid; i < G * YCR_MAX_LEVELS; i += K3_NT) s_rect[i] = ws.rect[b * G * YCR_MAX_LEVELS + i];
    if (tid == 0) { s_n = 0; s_nc = 0; }
    __syncthreads();
    for (int e = tid; e < G * topk; e += K3_NT) {
        const int g = e / topk;
        const int anchor = ws.sel[(int64_t)(b * G) * topk + e];
        if (anchor >= 0) atomicAdd(&s_a[anchor], 0x10000u + (uint32_t)g);
    }
    __syncthreads();
    int* pos_row = ws.pos_row + (int64_t)b * A;
    for (int an = tid; an < A; an += K3_NT) {
        const uint32_t v = s_a[an];
        const int cnt = v >> 16;
        pos_row[an] = -1;
        if (cnt == 1) {
            const int g = v & 0xFFFFu;
            const int idx = atomicAdd(&s_n, 1);
            if (idx < pos_cap) { s_pa[idx] = an; s_pg[idx] = g; }
            atomicAdd(&s_gcnt[g], 1);
        } else if (cnt > 1) {
            const int idx = atomicAdd(&s_nc, 1);
            if (idx < pos_cap) s_conf[idx] = an;
        }
    }
    __syncthreads();
    // anchors picked by several GTs: one warp per anchor, lanes over the image's GTs; the GT with the highest
    // overlap among ALL GTs whose box holds the anchor wins, the lowest index on ties (argmax), GT 0 if none
    {
        const int lane = tid & 31, wid = tid >> 5;
        const int nc = min(s_nc, pos_cap);
        for (int ci = wid; ci < nc; ci += K3_NT / 32) {
            const int an = s_conf[ci];
            const AnchorPos p = anchor_pos(a.grid, an);
            float best = 0.f;
            int g = 0;
            for (int gg = lane; gg < G; gg += 32) {
                if (s_goff[gg] < 0) continue;
                const int c = cand_index(a.grid, s_rect + gg * YCR_MAX_LEVELS, p);
                if (c < 0) continue;
                const float ov = ws.cand_ov[s_goff[gg] + c];
                if (ov > best) { best = ov; g = gg; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int og = __shfl_xor_sync(0xffffffffu, g, o);
                if (ob > best || (ob == best && og < g)) { best = ob; g = og; }
            }
            if (lane == 0) {
                const int idx = atomicAdd(&s_n, 1);
                if (idx < pos_cap) { s_pa[idx] = an; s_pg[idx] = g; }
                atomicAdd(&s_gcnt[g], 1);
            }
        }
    }
    __syncthreads();
    const int n = min(s_n, pos_cap);
    if (tid == 0) {
        int run = 0;
        for (int g = 0; g < G; ++g) { s_gstart[g] = run; run += s_gcnt[g]; }
    }
    __syncthreads();
    int* o_anchor = ws.pos_anchor + (int64_t)b * pos_cap;
    int* o_g = ws.pos_g + (int64_t)b * pos_cap;
    float* o_norm = ws.pos_norm + (int64_t)b * pos_cap;
    for (int p = tid; p < n; p += K3_NT) {
        const int an = s_pa[p], g = s_pg[p];
        const int64_t key = (int64_t)g * A + an;
        int rank = 0;
        for (int q = 0; q < n; ++q) rank += ((int64_t)s_pg[q] * A + s_pa[q] < key) ? 1 : 0;
        o_anchor[rank] = an;
        o_g[rank] = g;
        pos_row[an] = rank;
        float alv = 0.f, ovv = 0.f;
        const int ci = (s_goff[g] >= 0) ? cand_index(a.grid, s_rect + g * YCR_MAX_LEVELS, anchor_pos(a.grid, an)) : -1;
        if (ci >= 0) {
            alv = ws.cand_align[s_goff[g] + ci];
            ovv = ws.cand_ov[s_goff[g] + ci];
        }
        s_al[rank] = alv;
        atomicMax(&s_gal[g], __float_as_uint(fmaxf(alv, 0.f)));
        atomicMax(&s_gov[g], __float_as_uint(fmaxf(ovv, 0.f)));
    }
    __syncthreads();
    float part = 0.f;
    for (int r = tid; r < n; r += K3_NT) {
        const int g = o_g[r];
        const float norm = s_al[r] * __uint_as_float(s_gov[g]) / (__uint_as_float(s_gal[g]) + a.cfg.eps);
        o_norm[r] = norm;
        part += norm;
    }
    // fixed-shape tree: deterministic
    part = warp_sum(part);
    if ((tid & 31) == 0) s_red[tid >> 5] = part;
    __syncthreads();
    if (tid < 32) {
        float v = (tid < K3_NT / 32) ? s_red[tid] : 0.f;
        v = warp_sum(v);
        if (tid == 0) { ws.tss_part[b] = v; ws.npos[b] = n; }
    }
    for (int g = tid; g < G; g += K3_NT) {
        ws.gt_row_start[b * G + g] = s_gstart[g];
        ws.gt_row_cnt[b * G + g] = min(s_gcnt[g], max(0, pos_cap - s_gstart[g]));
    }
    // The last block to finish turns the per-image counts into row bases and the loss normaliser
    // target_scores_sum = max(sum, 1) (utils/loss.py:866), summed in image order (deterministic).
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(&ws.totals[3], 1) == (int)gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (s_last && tid < 32) {
        __threadfence();
        const int B = gridDim.x;
        int run = 0;
        double s = 0.0;
        for (int b0 = 0; b0 < B; b0 += 32) {
            const int bb = b0 + tid;
            const int np = (bb < B) ? __ldcg(&ws.npos[bb]) : 0;
            const double tp = (bb < B) ? (double)__ldcg(&ws.tss_part[bb]) : 0.0;
            int inc = np;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, inc, o);
                if (tid >= o) inc += v;
            }
            if (bb < B) ws.img_base[bb] = run + inc - np;
            run += __shfl_sync(0xffffffffu, inc, 31);
            for (int k = 0; k < min(32, B - b0); ++k) s += __shfl_sync(0xffffffffu, tp, k);
        }
        if (tid == 0) {
            ws.img_base[B] = run;
            ws.tss[0] = fmaxf((float)s, 1.f);
            ws.tss[1] = (float)s;
            if (n_pos_d) *n_pos_d = ws.err[0] ? -1 : run;  // -1: candidate capacity exceeded, nothing was assigned
            ws.totals[3] = 0;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K4: polar targets of the positives (utils/tal.py:1172-1193), centerness (utils/tal.py:1220), and,
// for the fused loss, the Polar-IoU log-ratio term of MaskIOULoss (utils/loss.py:113-127) with its
// gradient with respect to the raw ray outputs.  One block per GT.
// ------------------------------------------------------------------------------------------------
// Sum of R terms in the association order of k_positive_gather (lane i % 32 adds its terms in order, then
// an xor-butterfly 16..1 across the lanes), so that both forms of K4 give bit-identical sums.
template <int R>
__device__ __forceinline__ float butterfly_sum(const float (&term)[R]) {
    float p[32];
#pragma unroll
    for (int l = 0; l < 32; ++l) {
        p[l] = 0.f;
#pragma unroll
        for (int i = l; i < R; i += 32) p[l] += term[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int l = 0; l < o; ++l) p[l] += p[l + o];
    return p[0];
}

struct PosArgs {
    float* gt_dist; float* centerness; int pos_capacity;
    const int* img_base; const float* tss;
    int with_loss; float box_gain;
};

template <int R, int NT>
__global__ void __launch_bounds__(NT) k_positive_targets(const __grid_constant__ AssignArgs a, const __grid_constant__ AssignWs ws,
                                                          const PosArgs pa) {
    pdl_enter();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PolarSmem<R, NT>& sm = *reinterpret_cast<PolarSmem<R, NT>*>(smem_raw);
    const int bg = blockIdx.x, tid = threadIdx.x;
    const int cnt = ws.gt_row_cnt[bg];
    if (cnt == 0) return;
    const int b = bg / a.gt.G;
    init_raydir<R, NT>(sm, tid);
    {
        const float* cp = a.gt.coor + (int64_t)bg * a.gt.coor_stride;
        float* dst = reinterpret_cast<float*>(sm.contour);
        for (int k = tid; k < 2 * YCR_C; k += NT) dst[k] = cp[k];
    }
    const int row0 = ws.gt_row_start[bg];
    const int base = pa.img_base[b];
    for (int r0 = 0; r0 < cnt; r0 += NT) {
        __syncthreads();
        __syncthreads();
        const int r = r0 + tid;
        const bool active = r < cnt;
        const int row = row0 + r;
        AnchorPos ap{0, 0, 0, 0};
        float ax = 0.f, ay = 0.f;
        if (active) {
            ap = anchor_pos(a.grid, ws.pos_anchor[(int64_t)b * ws.pos_cap + row]);
            ax = anchor_coord(ap.ix, a.grid.stride[ap.level]);
            ay = anchor_coord(ap.iy, a.grid.stride[ap.level]);
            polar_sweep<R, NT>(sm, a.pc, tid, ax, ay);
        }
        const int nq = polar_settle_own<R, NT>(sm, a.pc, tid, active, ax, ay);
        polar_settle_queue<R, NT>(sm, a.pc, tid, nq);
        if (!active) continue;
        const int grow = base + row;
        float tmin = 3.4e38f, tmax = 0.f;
        for (int i = 0; i < R; ++i) {
            const float t = sm.tv(i, tid);
            tmin = fminf(tmin, t);
            tmax = fmaxf(tmax, t);
            if (pa.gt_dist && grow < pa.pos_capacity) pa.gt_dist[(int64_t)grow * R + i] = t;
        }
        if (pa.centerness && grow < pa.pos_capacity) pa.centerness[grow] = sqrtf(tmin / tmax);
        if (pa.with_loss) {
            const int l = ap.level;
            const int64_t r0 = (int64_t)b * a.pred.rays_sb[l] + (int64_t)ap.a_local * a.pred.rays_sa[l];
            const int64_t sc = a.pred.rays_sc[l];
            const float rs = a.pred.ray_scale[l];
            float mn[R], mx[R];
#pragma unroll
            for (int i = 0; i < R; ++i) {
                const float p = ycr_round_to(ycr_ld(a.pred.rays[l], r0 + i * sc, a.dtype) * rs, a.dtype);
                const float t = sm.tv(i, tid);
                mn[i] = fmaxf(fminf(p, t), YCR_FLOOR);
                mx[i] = fmaxf(p, t);
            }
            const float smin = butterfly_sum<R>(mn), smax = butterfly_sum<R>(mx);
            const float w = ws.pos_norm[(int64_t)b * ws.pos_cap + row];
            ws.pos_loss[(int64_t)b * ws.pos_cap + row] = logf(smax / smin) * w;
            const float coef = w / pa.tss[0] * pa.box_gain * (float)a.gt.B * rs;
            const float imax = 1.f / smax, imin = 1.f / smin;
            float* gp = ws.pos_grad + ((int64_t)b * ws.pos_cap + row) * R;
            for (int i = 0; i < R; ++i) {
                const float p = ycr_round_to(ycr_ld(a.pred.rays[l], r0 + i * sc, a.dtype) * rs, a.dtype);
                const float t = sm.tv(i, tid);
                float g = 0.f;
                if (p >= t) g += imax;                       // max() routes to pred (first index on ties)
                if (p <= t && p >= YCR_FLOOR) g -= imin;     // min() routes to pred; clamp passes when >= floor
                gp[i] = g * coef;
            }
        }
    }
}

// Ray target of (anchor, ray) straight from the contour in global memory, one thread, exact scan: only for a
// positive that lies outside its GT's box.  That happens when an anchor picked by several GTs has overlap 0 with
// all of them (predictions of +inf): argmax over all-zero overlaps is GT 0 (utils/tal.py:231), whatever its box,
// and the reference computes real geometry for every mask_pos entry (utils/tal.py:1172-1193).
template <int R>
__device__ __noinline__ float ray_target_global(const float* __restrict__ coor, const PolarConst& pc, float ax, float ay, int ray) {
    const double ang = (double)(ray * (360 / R)) * (3.14159265358979323846 / 180.0);
    const float cr = (float)cos(ang), sr = (float)sin(ang);
    uint4 K = make_uint4(YCR_EMPTY, YCR_EMPTY, YCR_EMPTY, YCR_EMPTY);
    for (int j = 0; j < YCR_C; ++j) {
        float vx = coor[2 * j] - ax;
        const float vy = coor[2 * j + 1] - ay;
        if (vx == 0.f && vy == 0.f) vx = 1.f;
        insert4(K.x, K.y, K.z, K.w, pack_pseudo(pseudo_angle(fabsf(fmaf(vy, cr, -vx * sr)), fmaf(vx, cr, vy * sr)), j));
    }
    if ((K.x >> 9) > pc.q2_gate) return YCR_FLOOR;
    const uint32_t e[4] = {K.x, K.y, K.z, K.w};
    float m = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int j = (int)(e[q] & 511u);
        const float vx = coor[2 * j] - ax, vy = coor[2 * j + 1] - ay;
        m = fmaxf(m, fmaf(vx, vx, vy * vy));
    }
    return fmaxf(dist_sqrt(m), YCR_FLOOR);
}

// K4 (gather form): when K1 kept the ray targets of every candidate, a positive's targets are just read
// back - one warp per positive, lanes over the rays; same outputs as k_positive_targets.
template <int R>
__global__ void __launch_bounds__(256) k_positive_gather(const __grid_constant__ AssignArgs a, const __grid_constant__ AssignWs ws,
                                                         const PosArgs pa) {
    pdl_enter();
    constexpr int NR = (R + 31) / 32;
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= ws.npos[b]) return;
    const int64_t prow = (int64_t)b * ws.pos_cap + row;
    const int an = ws.pos_anchor[prow];
    const int g = ws.pos_g[prow];
    const int bg = b * a.gt.G + g;
    const AnchorPos ap = anchor_pos(a.grid, an);
    const int l = ap.level;
    // the predictions do not depend on the candidate lookup: request them first
    const int64_t r0 = (int64_t)b * a.pred.rays_sb[l] + (int64_t)ap.a_local * a.pred.rays_sa[l];
    const int64_t sc = a.pred.rays_sc[l];
    const float rs = a.pred.ray_scale[l];
    float p[NR];
#pragma unroll
    for (int k = 0; k < NR; ++k) {
        const int i = lane + 32 * k;
        p[k] = (pa.with_loss && i < R) ? ycr_round_to(ycr_ld(a.pred.rays[l], r0 + i * sc, a.dtype) * rs, a.dtype) : 0.f;
    }
    const float w = ws.pos_norm[prow];
    const float tss = pa.tss[0];
    const int grow = pa.img_base[b] + row;
    const int ci = ws.valid[bg] ? cand_index(a.grid, ws.rect + bg * YCR_MAX_LEVELS, ap) : -1;
    const float* tp = nullptr;
    constexpr int NT1 = K1Nt<R>::value;
    if (ci >= 0) tp = ws.cand_t + ((int64_t)(ws.chunk_off[bg] + ci / NT1) * R) * NT1 + (ci % NT1);
    float t[NR];
    float tmin = 3.4e38f, tmax = 0.f, smin = 0.f, smax = 0.f;
#pragma unroll
    for (int k = 0; k < NR; ++k) {
        const int i = lane + 32 * k;
        if (i < R) {
            t[k] = tp ? tp[i * NT1]
                      : ray_target_global<R>(a.gt.coor + (int64_t)bg * a.gt.coor_stride, a.pc,
                                             anchor_coord(ap.ix, a.grid.stride[l]), anchor_coord(ap.iy, a.grid.stride[l]), i);
            tmin = fminf(tmin, t[k]);
            tmax = fmaxf(tmax, t[k]);
            smin += fmaxf(fminf(p[k], t[k]), YCR_FLOOR);
            smax += fmaxf(p[k], t[k]);
            if (pa.gt_dist && grow < pa.pos_capacity) pa.gt_dist[(int64_t)grow * R + i] = t[k];
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {   // fixed-shape butterflies: deterministic
        tmin = fminf(tmin, __shfl_xor_sync(0xffffffffu, tmin, o));
        tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
        smin += __shfl_xor_sync(0xffffffffu, smin, o);
        smax += __shfl_xor_sync(0xffffffffu, smax, o);
    }
    if (lane == 0 && pa.centerness && grow < pa.pos_capacity) pa.centerness[grow] = sqrtf(tmin / tmax);
    if (pa.with_loss) {
        if (lane == 0) ws.pos_loss[prow] = logf(smax / smin) * w;
        const float coef = w / tss * pa.box_gain * (float)a.gt.B * rs;
        const float imax = 1.f / smax, imin = 1.f / smin;
        float* gp = ws.pos_grad + prow * R;
#pragma unroll
        for (int k = 0; k < NR; ++k) {
            const int i = lane + 32 * k;
            if (i < R) {
                float gr = 0.f;
                if (p[k] >= t[k]) gr += imax;                       // max() routes to pred (first index on ties)
                if (p[k] <= t[k] && p[k] >= YCR_FLOOR) gr -= imin;  // min() routes to pred; clamp passes when >= floor
                gp[i] = gr * coef;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Dense API outputs of TaskAlignedAssigner.forward (get_targets utils/tal.py:1340-1390)
// ------------------------------------------------------------------------------------------------
__global__ void k_dense_targets(const __grid_constant__ AssignArgs a, const __grid_constant__ AssignWs ws,
                                const ycr_assign_out_t out) {
    const int A = a.grid.off[YCR_MAX_LEVELS];
    const int b = blockIdx.y;
    const int an = blockIdx.x * blockDim.x + threadIdx.x;
    if (an >= A) return;
    const int G = a.gt.G, nc = a.cfg.num_classes;
    const int row = ws.pos_row[(int64_t)b * A + an];
    const bool fg = row >= 0;
    const int g = fg ? ws.pos_g[(int64_t)b * ws.pos_cap + row] : 0;
    const int64_t i = (int64_t)b * A + an;
    const int bg = b * G + g;
    int64_t label = (int64_t)a.gt.labels[(int64_t)bg * a.gt.labels_stride];
    if (label < 0) label = 0;
    if (out.target_gt_idx_i64) out.target_gt_idx_i64[i] = g;
    if (out.fg_mask) out.fg_mask[i] = fg ? 1 : 0;
    if (out.target_labels_i64) out.target_labels_i64[i] = label;
    if (out.target_bboxes) {
        const float* bx = a.gt.boxes + (int64_t)bg * a.gt.boxes_stride;
        reinterpret_cast<float4*>(out.target_bboxes)[i] = make_float4(bx[0], bx[1], bx[2], bx[3]);
    }
    if (out.target_scores) {
        float* ts = out.target_scores + i * nc;
        for (int c = 0; c < nc; ++c) ts[c] = 0.f;
        if (fg && label < nc) ts[label] = ws.pos_norm[(int64_t)b * ws.pos_cap + row];
    }
    if (out.mask_pos && fg) out.mask_pos[((int64_t)b * G + g) * A + an] = 1;
}

__global__ void k_dense_metrics(const __grid_constant__ AssignArgs a, const __grid_constant__ AssignWs ws,
                                float* overlaps, float* align) {
    const int bg = blockIdx.x;
    const int A = a.grid.off[YCR_MAX_LEVELS];
    if (!ws.valid[bg] || ws.err[0]) return;
    const int n = ws.ncand[bg];
    const int4* rect = ws.rect + bg * YCR_MAX_LEVELS;
    for (int c = threadIdx.x; c < n; c += blockDim.x) {
        const AnchorPos p = cand_anchor(a.grid, rect, c);
        const int64_t o = (int64_t)bg * A + a.grid.off[p.level] + p.a_local;
        if (overlaps) overlaps[o] = ws.cand_ov[ws.cand_off[bg] + c];
        if (align) align[o] = ws.cand_align[ws.cand_off[bg] + c];
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
size_t assign_ws_layout(AssignWs* ws, void* base, const GridDev& grid, int B, int G, int topk, int R,
                        int64_t cand_cap, bool with_loss) {
    WsAlloc al{reinterpret_cast<char*>(base), 0, 0};
    const int BG = B * G, A = grid.off[YCR_MAX_LEVELS];
    const int pos_cap = (G * topk > 0) ? G * topk : 1;
    AssignWs w{};
    w.rect = al.take<int4>((size_t)BG * YCR_MAX_LEVELS + 1);
    w.ncand = al.take<int>(BG + 1);
    w.cand_off = al.take<int>(BG + 1);
    w.chunk_off = al.take<int>(BG + 1);
    const size_t nt1 = (size_t)k1_nt(R);
    const size_t chunks_cap = (size_t)cand_cap / nt1 + (size_t)BG + 1;  // every GT adds at most one partial chunk
    w.chunk_bg = al.take<int>(chunks_cap);
    w.chunk_work = al.take<int>(chunks_cap);
    w.part_off = al.take<int>(BG + 1);
    w.chunks_cap = (int)chunks_cap;
    w.valid = al.take<uint8_t>(BG + 1);
    w.totals = al.take<int>(4);
    w.err = al.take<int>(1);
    w.cand_align = al.take<float>((size_t)cand_cap + 1);
    w.cand_ov = al.take<float>((size_t)cand_cap + 1);
    {
        // keep the ray targets of every candidate when that is affordable (C2: 72 MB); the positives are then
        // a gather instead of a second sweep.  YCR_T_STORE_MAX_BYTES overrides the 1 GiB budget (0 = never).
        const char* env = getenv("YCR_T_STORE_MAX_BYTES");
        const size_t budget = env ? (size_t)strtoull(env, nullptr, 10) : ((size_t)1 << 30);
        const size_t bytes = chunks_cap * R * nt1 * sizeof(float);
        w.cand_t = (bytes <= budget && BG > 0) ? al.take<float>(chunks_cap * R * nt1) : nullptr;
    }
    w.sel = al.take<int>((size_t)BG * topk + 1);
    w.npos = al.take<int>(B + 1);
    w.pos_anchor = al.take<int>((size_t)B * pos_cap);
    w.pos_g = al.take<int>((size_t)B * pos_cap);
    w.pos_norm = al.take<float>((size_t)B * pos_cap);
    w.pos_row = al.take<int>((size_t)B * A);
    w.gt_row_start = al.take<int>(BG + 1);
    w.gt_row_cnt = al.take<int>(BG + 1);
    w.tss_part = al.take<float>(B + 1);
    w.img_base = al.take<int>(B + 2);
    w.tss = al.take<float>(2);
    if (with_loss) {
        w.pos_loss = al.take<float>((size_t)B * pos_cap);
        w.pos_grad = al.take<float>((size_t)B * pos_cap * R);
        w.n_bce_blocks = ((A + 255) / 256) * B;
        w.bce_part = al.take<float>((size_t)w.n_bce_blocks + 1);
    }
    w.pos_cap = pos_cap;
    w.cand_cap = cand_cap;
    if (ws) *ws = w;
    return align_up(al.off, 256);
}

template <int R>
static int launch_k1(const AssignArgs& a, const AssignWs& ws, cudaStream_t st) {
    constexpr int NT1 = K1Nt<R>::value;
    const size_t smem = sizeof(PolarSmem<R, NT1>);
    YCR_CUDA_CHECK(cudaFuncSetAttribute(k_cand_overlaps<R, NT1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    YCR_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_cand_overlaps<R, NT1>, NT1, smem));
    if (per_sm < 1) per_sm = 1;
    int dev = 0, sms = YCR_NUM_SMS;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    { YcrProfScope ps(YCR_T_CAND, st); YCR_CUDA_CHECK(ycr_launch(k_cand_overlaps<R, NT1>, dim3(sms * per_sm), dim3(NT1), smem, st, a, ws)); }
    YCR_LAUNCH_CHECK();
    return YCR_OK;
}

int launch_assign_core(const AssignArgs& a, const AssignWs& ws, int* n_pos_d, cudaStream_t st) {
    const int B = a.gt.B, G = a.gt.G, BG = B * G;
    const int A = a.grid.off[YCR_MAX_LEVELS];
    if (BG == 0) {  // otherwise k_gt_setup initialises both
        YCR_CUDA_CHECK(cudaMemsetAsync(ws.err, 0, sizeof(int), st));
        YCR_CUDA_CHECK(cudaMemsetAsync(ws.totals, 0, 4 * sizeof(int), st));
    }
    if (BG > 0) {
        {
            YcrProfScope ps(YCR_T_SETUP, st);
            YCR_CUDA_CHECK(ycr_launch(k_gt_rects, dim3((BG + 127) / 128), dim3(128), 0, st, a.grid, a.gt, ws));
            YCR_CUDA_CHECK(ycr_launch(k_gt_setup, dim3(1), dim3(1024), 0, st, a.gt, ws, k1_nt(a.cfg.rays)));
        }
        YCR_LAUNCH_CHECK();
        int rc = (a.cfg.rays == 36) ? launch_k1<36>(a, ws, st) : launch_k1<72>(a, ws, st);
        if (rc) return rc;
        { YcrProfScope ps(YCR_T_TOPK, st); YCR_CUDA_CHECK(ycr_launch(k_topk_per_gt, dim3((BG + 3) / 4), dim3(128), 0, st, a, ws)); }
        YCR_LAUNCH_CHECK();
    }
    const size_t smem3 = (size_t)ycr_resolve_smem_bytes(A, G, a.cfg.topk);
    YCR_CUDA_CHECK(cudaFuncSetAttribute(k_resolve_image, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3));
    { YcrProfScope ps(YCR_T_RESOLVE, st); YCR_CUDA_CHECK(ycr_launch(k_resolve_image, dim3(B), dim3(K3_NT), smem3, st, a, ws, n_pos_d)); }
    YCR_LAUNCH_CHECK();
    return YCR_OK;
}

int launch_positive_targets(const AssignArgs& a, const AssignWs& ws, float* gt_dist, float* centerness,
                            int pos_capacity, bool with_loss, const ycr_loss_cfg_t* lcfg,
                            cudaStream_t st) {
    const int B = a.gt.B, BG = B * a.gt.G;
    int* img_base = ws.img_base;
    float* tss = ws.tss;
    PosArgs pa{gt_dist, centerness, pos_capacity, img_base, tss, with_loss ? 1 : 0, lcfg ? lcfg->box_gain : 0.f};
    if (BG == 0) return YCR_OK;
    YcrProfScope ps(YCR_T_POS, st);
    if (ws.cand_t) {
        dim3 grid((ws.pos_cap + 7) / 8, B);
        if (a.cfg.rays == 36) YCR_CUDA_CHECK(ycr_launch(k_positive_gather<36>, grid, dim3(256), 0, st, a, ws, pa));
        else YCR_CUDA_CHECK(ycr_launch(k_positive_gather<72>, grid, dim3(256), 0, st, a, ws, pa));
        YCR_LAUNCH_CHECK();
        return YCR_OK;
    }
    if (a.cfg.rays == 36) {
        const size_t smem = sizeof(PolarSmem<36, K4_NT>);
        YCR_CUDA_CHECK(cudaFuncSetAttribute(k_positive_targets<36, K4_NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        YCR_CUDA_CHECK(ycr_launch(k_positive_targets<36, K4_NT>, dim3(BG), dim3(K4_NT), smem, st, a, ws, pa));
    } else {
        const size_t smem = sizeof(PolarSmem<72, K4_NT>);
        YCR_CUDA_CHECK(cudaFuncSetAttribute(k_positive_targets<72, K4_NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        YCR_CUDA_CHECK(ycr_launch(k_positive_targets<72, K4_NT>, dim3(BG), dim3(K4_NT), smem, st, a, ws, pa));
    }
    YCR_LAUNCH_CHECK();
    return YCR_OK;
}

int launch_assign_dense(const AssignArgs& a, const AssignWs& ws, const ycr_assign_out_t& out, cudaStream_t st) {
    const int B = a.gt.B, G = a.gt.G, BG = B * G;
    const int A = a.grid.off[YCR_MAX_LEVELS];
    if (out.mask_pos) YCR_CUDA_CHECK(cudaMemsetAsync(out.mask_pos, 0, (size_t)BG * A, st));
    dim3 grid((A + 255) / 256, B);
    k_dense_targets<<<grid, 256, 0, st>>>(a, ws, out);
    YCR_LAUNCH_CHECK();
    if (out.overlaps || out.align_metric) {
        if (out.overlaps) YCR_CUDA_CHECK(cudaMemsetAsync(out.overlaps, 0, (size_t)BG * A * 4, st));
        if (out.align_metric) YCR_CUDA_CHECK(cudaMemsetAsync(out.align_metric, 0, (size_t)BG * A * 4, st));
        k_dense_metrics<<<BG, 128, 0, st>>>(a, ws, out.overlaps, out.align_metric);
        YCR_LAUNCH_CHECK();
    }
    return YCR_OK;
}
