// Inference path: Segment head polar decode (nn/modules/head.py:461-494, 559-570) and
// non_max_suppression, polar variant (utils/ops.py:285-424) with the torchvision.ops.nms step
// (utils/ops.py:407) as a sorted, chunked block-bitmask greedy suppression.
// Paths under /root/reference/ultralytics-main/ultralytics/.
#include "common.cuh"
#include "dtype.cuh"
#include <math.h>

// ------------------------------------------------------------------------------------------------
// decode: one thread per (image, anchor); every access of a warp is a contiguous line along the
// anchor dimension of the channel-major input (B, R+nc, HW_l) and output (B, 4+nc+3R, A).
// Arithmetic mirrors the reference op by op (separate multiply and add, clamp after the stride
// multiply) so boxes agree to the last bit wherever sin/cos tables do.
// ------------------------------------------------------------------------------------------------
struct DecodeArgs {
    GridDev grid;
    const void* feats[YCR_MAX_LEVELS];   // element type `dtype` (fp32 / fp16 / bf16); arithmetic and output are fp32
    int B, nc, R, dtype;
    float cs[2 * 72];  // cos[0..R), sin[0..R)
};

__global__ void __launch_bounds__(256) k_decode(const __grid_constant__ DecodeArgs d, float* __restrict__ out) {
    const int A = d.grid.off[YCR_MAX_LEVELS];
    const int b = blockIdx.y;
    const int an = blockIdx.x * 256 + threadIdx.x;
    if (an >= A) return;
    int l = 0;
#pragma unroll
    for (int k = 1; k < YCR_MAX_LEVELS; ++k)
        if (k < d.grid.n_levels && an >= d.grid.off[k]) l = k;
    const int hw = d.grid.h[l] * d.grid.w[l];
    const int al = an - d.grid.off[l];
    const int iy = al / d.grid.w[l], ix = al - iy * d.grid.w[l];
    const float stride = d.grid.stride[l];
    const float ax = ((float)ix + 0.5f) * stride, ay = ((float)iy + 0.5f) * stride;
    const int R = d.R, nc = d.nc;
    const int CH = 4 + nc + 3 * R;
    const int64_t f0 = (int64_t)b * (R + nc) * hw + al;
    float* o = out + (int64_t)b * CH * A + an;
    float minx = 3.4e38f, miny = 3.4e38f, maxx = -3.4e38f, maxy = -3.4e38f;
#pragma unroll 4
    for (int i = 0; i < R; ++i) {
        const float dist = fmaxf(__fmul_rn(ycr_ld(d.feats[l], f0 + (int64_t)i * hw, d.dtype), stride), YCR_FLOOR);
        const float x = __fadd_rn(__fmul_rn(dist, d.cs[i]), ax);
        const float y = __fadd_rn(__fmul_rn(dist, d.cs[R + i]), ay);
        minx = fminf(minx, x); maxx = fmaxf(maxx, x);
        miny = fminf(miny, y); maxy = fmaxf(maxy, y);
        o[(int64_t)(4 + nc + i) * A] = x;
        o[(int64_t)(4 + nc + R + i) * A] = y;
        o[(int64_t)(4 + nc + 2 * R + i) * A] = (dist > 1.f) ? 1.f : 0.f;
    }
    o[0] = minx;
    o[(int64_t)A] = miny;
    o[(int64_t)2 * A] = maxx;
    o[(int64_t)3 * A] = maxy;
#pragma unroll 4
    for (int c = 0; c < nc; ++c) {
        const float x = ycr_ld(d.feats[l], f0 + (int64_t)(R + c) * hw, d.dtype);
        o[(int64_t)(4 + c) * A] = 1.f / (1.f + expf(-x));
    }
}

// Vectorised variant: one thread per four consecutive anchors of a row (W_l and H_l*W_l multiples of 4),
// 128-bit loads/stores.  Split in two kernels so that the class part (pure streaming sigmoid) is not
// held to the occupancy of the ray part (which carries the 16 running box extrema).
template <typename T>
__global__ void __launch_bounds__(256, 3) k_decode_rays_v4(const __grid_constant__ DecodeArgs d, float* __restrict__ out) {
    const int A = d.grid.off[YCR_MAX_LEVELS];
    const int b = blockIdx.y;
    const int an = (blockIdx.x * 256 + threadIdx.x) * 4;
    if (an >= A) return;
    int l = 0;
#pragma unroll
    for (int k = 1; k < YCR_MAX_LEVELS; ++k)
        if (k < d.grid.n_levels && an >= d.grid.off[k]) l = k;
    const int hw = d.grid.h[l] * d.grid.w[l];
    const int al = an - d.grid.off[l];
    const int iy = al / d.grid.w[l], ix = al - iy * d.grid.w[l];
    const float stride = d.grid.stride[l];
    const float ay = ((float)iy + 0.5f) * stride;
    const float ax0 = ((float)ix + 0.5f) * stride, ax1 = ((float)(ix + 1) + 0.5f) * stride;
    const float ax2 = ((float)(ix + 2) + 0.5f) * stride, ax3 = ((float)(ix + 3) + 0.5f) * stride;
    const int R = d.R, nc = d.nc;
    const int CH = 4 + nc + 3 * R;
    const T* f = reinterpret_cast<const T*>(d.feats[l]) + (int64_t)b * (R + nc) * hw + al;
    float* o = out + (int64_t)b * CH * A + an;
    float4 minx = make_float4(3.4e38f, 3.4e38f, 3.4e38f, 3.4e38f), miny = minx;
    float4 maxx = make_float4(-3.4e38f, -3.4e38f, -3.4e38f, -3.4e38f), maxy = maxx;
#pragma unroll 2
    for (int i = 0; i < R; ++i) {
        const float4 r = YcrType<T>::ld4cs(f + (int64_t)i * hw);
        const float c = d.cs[i], s = d.cs[R + i];
        float4 dist, x, y, v;
        dist.x = fmaxf(__fmul_rn(r.x, stride), YCR_FLOOR); dist.y = fmaxf(__fmul_rn(r.y, stride), YCR_FLOOR);
        dist.z = fmaxf(__fmul_rn(r.z, stride), YCR_FLOOR); dist.w = fmaxf(__fmul_rn(r.w, stride), YCR_FLOOR);
        x.x = __fadd_rn(__fmul_rn(dist.x, c), ax0); x.y = __fadd_rn(__fmul_rn(dist.y, c), ax1);
        x.z = __fadd_rn(__fmul_rn(dist.z, c), ax2); x.w = __fadd_rn(__fmul_rn(dist.w, c), ax3);
        y.x = __fadd_rn(__fmul_rn(dist.x, s), ay); y.y = __fadd_rn(__fmul_rn(dist.y, s), ay);
        y.z = __fadd_rn(__fmul_rn(dist.z, s), ay); y.w = __fadd_rn(__fmul_rn(dist.w, s), ay);
        v.x = (dist.x > 1.f) ? 1.f : 0.f; v.y = (dist.y > 1.f) ? 1.f : 0.f;
        v.z = (dist.z > 1.f) ? 1.f : 0.f; v.w = (dist.w > 1.f) ? 1.f : 0.f;
        minx.x = fminf(minx.x, x.x); minx.y = fminf(minx.y, x.y); minx.z = fminf(minx.z, x.z); minx.w = fminf(minx.w, x.w);
        maxx.x = fmaxf(maxx.x, x.x); maxx.y = fmaxf(maxx.y, x.y); maxx.z = fmaxf(maxx.z, x.z); maxx.w = fmaxf(maxx.w, x.w);
        miny.x = fminf(miny.x, y.x); miny.y = fminf(miny.y, y.y); miny.z = fminf(miny.z, y.z); miny.w = fminf(miny.w, y.w);
        maxy.x = fmaxf(maxy.x, y.x); maxy.y = fmaxf(maxy.y, y.y); maxy.z = fmaxf(maxy.z, y.z); maxy.w = fmaxf(maxy.w, y.w);
        __stcs(reinterpret_cast<float4*>(o + (int64_t)(4 + nc + i) * A), x);
        __stcs(reinterpret_cast<float4*>(o + (int64_t)(4 + nc + R + i) * A), y);
        __stcs(reinterpret_cast<float4*>(o + (int64_t)(4 + nc + 2 * R + i) * A), v);
    }
    *reinterpret_cast<float4*>(o) = minx;
    *reinterpret_cast<float4*>(o + (int64_t)A) = miny;
    *reinterpret_cast<float4*>(o + (int64_t)2 * A) = maxx;
    *reinterpret_cast<float4*>(o + (int64_t)3 * A) = maxy;
}

// class rows: grid (anchor groups, class groups of 8, B)
template <typename T>
__global__ void __launch_bounds__(256) k_decode_cls_v4(const __grid_constant__ DecodeArgs d, float* __restrict__ out) {
    const int A = d.grid.off[YCR_MAX_LEVELS];
    const int b = blockIdx.z;
    const int an = (blockIdx.x * 256 + threadIdx.x) * 4;
    if (an >= A) return;
    int l = 0;
#pragma unroll
    for (int k = 1; k < YCR_MAX_LEVELS; ++k)
        if (k < d.grid.n_levels && an >= d.grid.off[k]) l = k;
    const int hw = d.grid.h[l] * d.grid.w[l];
    const int al = an - d.grid.off[l];
    const int R = d.R, nc = d.nc;
    const int CH = 4 + nc + 3 * R;
    const int c0 = blockIdx.y * 8, c1 = min(nc, c0 + 8);
    const T* fc = reinterpret_cast<const T*>(d.feats[l]) + (int64_t)b * (R + nc) * hw + (int64_t)R * hw + al;
    float* o = out + (int64_t)b * CH * A + an;
#pragma unroll 8
    for (int c = c0; c < c1; ++c) {
        const float4 x = YcrType<T>::ld4cs(fc + (int64_t)c * hw);
        float4 p;
        p.x = 1.f / (1.f + expf(-x.x)); p.y = 1.f / (1.f + expf(-x.y));
        p.z = 1.f / (1.f + expf(-x.z)); p.w = 1.f / (1.f + expf(-x.w));
        *reinterpret_cast<float4*>(o + (int64_t)(4 + c) * A) = p;   // class rows are re-read by NMS
    }
}

// class rows and, per anchor, the best class {score, class index} (first maximum) that single-label NMS
// starts from (`conf, j = cls.max(1)`, utils/ops.py:386): one thread per four anchors loops over all classes,
// so the maximum needs no cross-thread step and NMS does not have to read the class rows back.
// (measured: unroll 1 with 8 blocks/SM - 32 registers - 0.51 ms for decode at C3; unroll 4 / 4 blocks 0.60 ms)
#ifndef DCB_MINB
#define DCB_MINB 8
#endif
#ifndef DCB_UNROLL
#define DCB_UNROLL 1
#endif
#define DCB_STR(x) #x
#define DCB_PRAGMA(n) _Pragma(DCB_STR(unroll n))
template <typename T>
__global__ void __launch_bounds__(256, DCB_MINB) k_decode_cls_best_v4(const __grid_constant__ DecodeArgs d, float* __restrict__ out,
                                                               int2* __restrict__ best) {
    const int A = d.grid.off[YCR_MAX_LEVELS];
    const int b = blockIdx.y;
    const int an = (blockIdx.x * 256 + threadIdx.x) * 4;
    if (an >= A) return;
    int l = 0;
#pragma unroll
    for (int k = 1; k < YCR_MAX_LEVELS; ++k)
        if (k < d.grid.n_levels && an >= d.grid.off[k]) l = k;
    const int hw = d.grid.h[l] * d.grid.w[l];
    const int al = an - d.grid.off[l];
    const int R = d.R, nc = d.nc;
    const int CH = 4 + nc + 3 * R;
    const T* fc = reinterpret_cast<const T*>(d.feats[l]) + (int64_t)b * (R + nc) * hw + (int64_t)R * hw + al;
    float* o = out + (int64_t)b * CH * A + an;
    float bs[4] = {-3.4e38f, -3.4e38f, -3.4e38f, -3.4e38f};
    int bc[4] = {0, 0, 0, 0};
DCB_PRAGMA(DCB_UNROLL)
    for (int c = 0; c < nc; ++c) {
        const float4 x = YcrType<T>::ld4cs(fc + (int64_t)c * hw);
        float4 p;
        p.x = 1.f / (1.f + expf(-x.x)); p.y = 1.f / (1.f + expf(-x.y));
        p.z = 1.f / (1.f + expf(-x.z)); p.w = 1.f / (1.f + expf(-x.w));
        *reinterpret_cast<float4*>(o + (int64_t)(4 + c) * A) = p;
        if (p.x > bs[0]) { bs[0] = p.x; bc[0] = c; }
        if (p.y > bs[1]) { bs[1] = p.y; bc[1] = c; }
        if (p.z > bs[2]) { bs[2] = p.z; bc[2] = c; }
        if (p.w > bs[3]) { bs[3] = p.w; bc[3] = c; }
    }
    int2* bo = best + (int64_t)b * A + an;
    reinterpret_cast<int4*>(bo)[0] = make_int4(__float_as_int(bs[0]), bc[0], __float_as_int(bs[1]), bc[1]);
    reinterpret_cast<int4*>(bo)[1] = make_int4(__float_as_int(bs[2]), bc[2], __float_as_int(bs[3]), bc[3]);
}

template <typename T>
static void launch_decode_v4(const DecodeArgs& d, dim3 g, dim3 gc, float* allpred, int2* best, cudaStream_t st) {
    k_decode_rays_v4<T><<<g, 256, 0, st>>>(d, allpred);
    if (best && reinterpret_cast<uintptr_t>(best) % 16 == 0) k_decode_cls_best_v4<T><<<g, 256, 0, st>>>(d, allpred, best);
    else k_decode_cls_v4<T><<<gc, 256, 0, st>>>(d, allpred);
}

int launch_decode(const ycr_grid_t* grid, const void* const* feats, int dtype, int B, int nc, int R, float* allpred, int2* best,
                  cudaStream_t st) {
    DecodeArgs d{};
    d.grid = make_grid_dev(grid);
    for (int l = 0; l < grid->n_levels; ++l) d.feats[l] = feats[l];
    d.B = B; d.nc = nc; d.R = R; d.dtype = dtype;
    for (int i = 0; i < R; ++i) {
        // angles = arange(0,360,360//R)/180.*pi in fp32 (nn/modules/head.py:466), then sin/cos
        const float deg = (float)(i * (360 / R));
        const float ang = (deg / 180.f) * (float)3.141592653589793;
        d.cs[i] = (float)cos((double)ang);
        d.cs[R + i] = (float)sin((double)ang);
    }
    const int A = d.grid.off[YCR_MAX_LEVELS];
    const uintptr_t align = 4 * ycr_dtype_size(dtype);
    bool vec = (reinterpret_cast<uintptr_t>(allpred) % 16 == 0);
    for (int l = 0; l < grid->n_levels; ++l)
        vec = vec && (grid->w[l] % 4 == 0) && (reinterpret_cast<uintptr_t>(feats[l]) % align == 0);
    if (vec) {
        dim3 g((A / 4 + 255) / 256, B);
        dim3 gc((A / 4 + 255) / 256, (nc + 7) / 8, B);
        YcrProfScope ps(YCR_T_DECODE, st);
        if (dtype == YCR_F16) launch_decode_v4<__half>(d, g, gc, allpred, best, st);
        else if (dtype == YCR_BF16) launch_decode_v4<__nv_bfloat16>(d, g, gc, allpred, best, st);
        else launch_decode_v4<float>(d, g, gc, allpred, best, st);
    } else {
        dim3 g((A + 255) / 256, B);
        YcrProfScope ps(YCR_T_DECODE, st);
        k_decode<<<g, 256, 0, st>>>(d, allpred);
        if (best) YCR_CUDA_CHECK(cudaMemsetAsync(best, 0xFF, (size_t)B * A * sizeof(int2), st));  // class -1: no hint
    }
    YCR_LAUNCH_CHECK();
    return YCR_OK;
}

// ------------------------------------------------------------------------------------------------
// NMS
// ------------------------------------------------------------------------------------------------
struct NmsWs {
    unsigned long long* keys;  // [B][cap2]  (~score bits << 32) | (anchor*nc + class), sorted ascending
    int* count;                // [B] candidates written (may exceed cap)
    float4* boxes;             // [B][nsel_cap] class-offset boxes in sorted order
    int4* kept;                // [B][NMS_MAX_KEEP] (anchor, class, score bits, -) of the kept rows
    unsigned long long* keys2;  // [B][presel_cap2] the preselected candidates (only when cap > NMS_PRESEL_MIN)
    unsigned long long* tkeys;  // [B][NMS_SORT_SMEM] the sorted leading tranche (only when cap > NMS_TRANCHE)
    int* nsorted;              // [B] entries of the sorted list the suppression may walk
    int* tmode;                // [B] 1: that list is the tranche in tkeys and more candidates exist behind it
    int* redo;                 // [B] 1: the tranche ran out before max_det boxes were kept -> second pass over everything
    // score-histogram preselection (multi-label with a very low threshold: hundreds of thousands of pairs per image pass)
    int* hist;                 // [B][NMS_HBINS] entries per score bin, best scores first (null: path off)
    int* bsel;                 // [B] last bin the first filter pass lets through
    int* more;                 // [B] entries that pass conf_thres but were left out by that bin limit
    int* row_base;             // [B] first output row of image b in the compact layout (exclusive scan of the kept counts)
    int* done_ctr;             // blocks of the final suppression launch that have finished
    int cap, cap2, nsel_cap, presel_cap2;
};
#define NMS_HBINS 2048

#define NMS_SORT_SMEM 4096
#define NMS_PRESEL_MIN 32768   // above this many candidates the top max_nms are selected before sorting
#define NMS_TRANCHE 1024       // candidates sorted first; greedy NMS rarely looks further before max_det are kept

static inline int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

static size_t nms_ws_layout(NmsWs* ws, void* base, int B, int A, const ycr_nms_cfg_t* cfg) {
    WsAlloc al{reinterpret_cast<char*>(base), 0, 0};
    NmsWs w{};
    const int nc = cfg->nc;
    const bool multi = cfg->multi_label && nc > 1;
    w.cap = multi ? A * nc : A;
    w.cap2 = next_pow2(w.cap);
    w.nsel_cap = (w.cap < cfg->max_nms) ? w.cap : cfg->max_nms;
    w.keys = al.take<unsigned long long>((size_t)B * w.cap2);
    w.count = al.take<int>(B + 1);
    w.boxes = al.take<float4>((size_t)B * w.nsel_cap + 1);
    w.kept = al.take<int4>((size_t)B * 1024);
    w.nsorted = al.take<int>(B);
    w.tmode = al.take<int>(B);
    w.redo = al.take<int>(B + 1);
    w.row_base = al.take<int>(B);
    w.done_ctr = al.take<int>(4);
    w.tkeys = (w.cap > NMS_TRANCHE) ? al.take<unsigned long long>((size_t)B * NMS_SORT_SMEM) : nullptr;
    w.hist = nullptr; w.bsel = nullptr; w.more = nullptr;
    if (multi && w.cap > NMS_PRESEL_MIN) {
        w.hist = al.take<int>((size_t)B * NMS_HBINS);
        w.bsel = al.take<int>(B);
        w.more = al.take<int>(B);
    }
    w.presel_cap2 = 0;
    w.keys2 = nullptr;
    if (w.cap > 32768 && w.cap > w.nsel_cap) {
        w.presel_cap2 = next_pow2(2 * w.nsel_cap);
        w.keys2 = al.take<unsigned long long>((size_t)B * w.presel_cap2);
    }
    if (ws) *ws = w;
    return align_up(al.off, 256);
}

// The head feature maps a prediction tensor was decoded from: lets the NMS kernels recompute what they need of an
// anchor (its box, its contour row) from the R ray values instead of reading the channel-major prediction.
struct GatherFeats {
    GridDev grid;
    const void* feats[YCR_MAX_LEVELS];
    int dtype, R;
    float cs[2 * 72];
};

// box (min x, min y, max x, max y) of anchor `an` of image b, arithmetic identical to k_decode*
__device__ __forceinline__ float4 box_from_feats(const GatherFeats& gf, int b, int an, int nc) {
    int l = 0;
#pragma unroll
    for (int q = 1; q < YCR_MAX_LEVELS; ++q)
        if (q < gf.grid.n_levels && an >= gf.grid.off[q]) l = q;
    const int hw = gf.grid.h[l] * gf.grid.w[l];
    const int al = an - gf.grid.off[l];
    const int iy = al / gf.grid.w[l], ix = al - iy * gf.grid.w[l];
    const float stride = gf.grid.stride[l];
    const float ax = ((float)ix + 0.5f) * stride, ay = ((float)iy + 0.5f) * stride;
    const int64_t f0 = (int64_t)b * (gf.R + nc) * hw + al;
    float minx = 3.4e38f, miny = 3.4e38f, maxx = -3.4e38f, maxy = -3.4e38f;
    // twelve strided loads in flight at a time (R is 36 or 72), raw words first: a conversion right behind each load
    // would serialise their latencies
    for (int i0 = 0; i0 < gf.R; i0 += 12) {
        uint32_t raw[12];
        if (gf.dtype == YCR_F32) {
            const uint32_t* p = reinterpret_cast<const uint32_t*>(gf.feats[l]) + f0 + (int64_t)i0 * hw;
#pragma unroll
            for (int k = 0; k < 12; ++k) raw[k] = (i0 + k < gf.R) ? p[(int64_t)k * hw] : 0u;
        } else {
            const unsigned short* p = reinterpret_cast<const unsigned short*>(gf.feats[l]) + f0 + (int64_t)i0 * hw;
#pragma unroll
            for (int k = 0; k < 12; ++k) raw[k] = (i0 + k < gf.R) ? p[(int64_t)k * hw] : 0u;
        }
#pragma unroll
        for (int k = 0; k < 12; ++k) {
            const int i = i0 + k;
            if (i < gf.R) {
                const float dist = fmaxf(__fmul_rn(ycr_from_raw(raw[k], gf.dtype), stride), YCR_FLOOR);
                const float x = __fadd_rn(__fmul_rn(dist, gf.cs[i]), ax);
                const float y = __fadd_rn(__fmul_rn(dist, gf.cs[gf.R + i]), ay);
                minx = fminf(minx, x); maxx = fmaxf(maxx, x);
                miny = fminf(miny, y); maxy = fmaxf(maxy, y);
            }
        }
    }
    return make_float4(minx, miny, maxx, maxy);
}

// Score bin of a confidence in (0, 1]: 2^16 float steps per bin counted down from 1.0 (bin 0 = the best scores;
// monotone, so "bin <= k" is a score threshold).  Everything below ~2^-63 lands in the last bin.
__device__ __forceinline__ int nms_score_bin(float v) {
    const int d = (int)(0x3F800000u - __float_as_uint(v));
    return min(max(d >> 16, 0), NMS_HBINS - 1);
}

__device__ __forceinline__ bool nms_class_ok(const ycr_nms_cfg_t& cfg, int c) {
    if (!cfg.classes) return true;
    for (int k = 0; k < cfg.n_classes; ++k)
        if (cfg.classes[k] == c) return true;
    return false;
}

// Multi-label NMS with a very low threshold (the validator: conf 0.001) lets almost every (anchor, class) pair through -
// 6*10^5 per image at 8400 anchors x 80 classes - while greedy NMS stops after max_det survivors, a few hundred entries
// down the sorted list.  So the pairs are first only COUNTED per score bin (this kernel), k_nms_pick turns the counts
// into a per-image bin limit that lets roughly NMS_TRANCHE pairs through, and the filter writes just those.  If the
// suppression runs out of entries before max_det boxes are kept, the image is redone with the plain filter.
__global__ void __launch_bounds__(256) k_nms_hist(const float* __restrict__ pred, int CH, int A, ycr_nms_cfg_t cfg, NmsWs ws) {
    __shared__ int s_h[NMS_HBINS];
    const int b = blockIdx.y;
    for (int k = threadIdx.x; k < NMS_HBINS; k += 256) s_h[k] = 0;
    __syncthreads();
    const int an = blockIdx.x * 256 + threadIdx.x;
    if (an < A) {
        const float* p = pred + ((int64_t)b * CH + 4) * A + an;
        for (int c = 0; c < cfg.nc; ++c) {
            const float v = p[(int64_t)c * A];
            if (v > cfg.conf_thres && nms_class_ok(cfg, c)) atomicAdd(&s_h[nms_score_bin(v)], 1);
        }
    }
    __syncthreads();
    int* h = ws.hist + (int64_t)b * NMS_HBINS;
    for (int k = threadIdx.x; k < NMS_HBINS; k += 256)
        if (s_h[k]) atomicAdd(&h[k], s_h[k]);
}

#define NMS_PICK (NMS_TRANCHE * 3 / 4)   // (a bin holds a few hundred entries at most: the pass usually stays below NMS_TRANCHE)
// per image: the first bin at which the running count reaches NMS_PICK; everything, if the image has fewer than
// NMS_SORT_SMEM entries anyway.  Also zeroes the image's entry counter for the filter that follows.
__global__ void __launch_bounds__(256) k_nms_pick(NmsWs ws) {
    __shared__ int s_part[256];
    __shared__ int s_sel[2];
    const int b = blockIdx.x, t = threadIdx.x;
    const int* h = ws.hist + (int64_t)b * NMS_HBINS;
    constexpr int PER = NMS_HBINS / 256;
    int local = 0;
#pragma unroll
    for (int k = 0; k < PER; ++k) local += h[t * PER + k];
    s_part[t] = local;
    if (t == 0) { s_sel[0] = NMS_HBINS - 1; s_sel[1] = -1; }
    __syncthreads();
    for (int o = 1; o < 256; o <<= 1) {   // inclusive scan of the 256 partial sums
        const int v = (t >= o) ? s_part[t - o] : 0;
        __syncthreads();
        s_part[t] += v;
        __syncthreads();
    }
    const int total = s_part[255];
    int cum = s_part[t] - local;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const int c = h[t * PER + k];
        if (cum < NMS_PICK && cum + c >= NMS_PICK) { s_sel[0] = t * PER + k; s_sel[1] = cum + c; }
        cum += c;
    }
    __syncthreads();
    if (t == 0) {
        const bool all = total <= NMS_SORT_SMEM || s_sel[1] < 0;
        ws.bsel[b] = all ? NMS_HBINS - 1 : s_sel[0];
        ws.more[b] = all ? 0 : total - s_sel[1];
        ws.count[b] = 0;
    }
}

// conf filter + best-class / multi-label expansion (utils/ops.py:348, 380-391)
// bin_limit: the per-image score-bin limits of k_nms_pick (null = everything above conf_thres); redo_only: only the
// images whose first, limited pass ran out of entries (ws.redo), this time without the limit.
__global__ void __launch_bounds__(256) k_nms_filter(const float* __restrict__ pred, int CH, int A, ycr_nms_cfg_t cfg, NmsWs ws,
                                                    const int* __restrict__ bin_limit, int redo_only) {
    const int b = blockIdx.y;
    if (redo_only && !(ws.redo[b] && ws.more && ws.more[b] > 0)) return;
    const int blim = bin_limit ? bin_limit[b] : NMS_HBINS;
    const int an = blockIdx.x * 256 + threadIdx.x;
    const unsigned lane = threadIdx.x & 31u;
    const bool live = an < A;
    const int nc = cfg.nc;
    const bool multi = cfg.multi_label && nc > 1;
    const float* p = pred + ((int64_t)b * CH + 4) * A + (live ? an : 0);
    float best = -3.4e38f;
    int bc = 0, npass = 0;
    const int2* hint = reinterpret_cast<const int2*>(cfg.best_class);
    int2 hv = make_int2(0, -1);
    if (live && hint && !multi) hv = hint[(int64_t)b * A + an];
    if (hv.y >= 0) {   // the decode kernel already found the best class of this anchor
        best = __int_as_float(hv.x);
        bc = hv.y;
    } else if (live) {
        for (int c = 0; c < nc; ++c) {
            const float v = p[(int64_t)c * A];
            if (v > best) { best = v; bc = c; }  // first maximum
            // (with a bin limit the entries of filtered-out classes are left out, as k_nms_hist counted them)
            npass += (v > cfg.conf_thres && (!bin_limit || (nms_score_bin(v) <= blim && nms_class_ok(cfg, c)))) ? 1 : 0;
        }
    }
    auto class_ok = [&](int c) { return nms_class_ok(cfg, c); };
    // entries this anchor writes; the slots of a warp are reserved with ONE atomic on the image's counter (the
    // validator's conf 0.001 lets nearly every anchor through: one atomic per anchor serialised the whole kernel)
    const bool pass = live && (best > cfg.conf_thres);
    int mine = 0;
    if (pass) mine = multi ? npass : (class_ok(bc) ? 1 : 0);
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (total == 0) return;   // (warp-uniform)
    int base = 0;
    if (lane == 31) base = atomicAdd(&ws.count[b], total);
    base = __shfl_sync(0xffffffffu, base, 31);
    if (mine == 0) return;
    int slot = base + incl - mine;
    unsigned long long* keys = ws.keys + (int64_t)b * ws.cap2;
    if (!multi) {
        if (slot < ws.cap)
            keys[slot] = ((unsigned long long)(~__float_as_uint(best)) << 32) | (unsigned)(an * nc + bc);
    } else {
        for (int c = 0; c < nc; ++c) {
            const float v = p[(int64_t)c * A];
            if (v > cfg.conf_thres && (!bin_limit || (nms_score_bin(v) <= blim && nms_class_ok(cfg, c)))) {
                // entries of filtered classes keep their slot but sort to the end and are cut off
                const bool ok = class_ok(c);
                if (slot < ws.cap)
                    keys[slot] = ok ? (((unsigned long long)(~__float_as_uint(v)) << 32) | (unsigned)(an * nc + c))
                                    : 0xFFFFFFFFFFFFFFFFull;
                ++slot;
            }
        }
    }
}

// block-wide bitonic sort of npow2 keys (shared or global memory)
__device__ void bitonic_sort(unsigned long long* d, int npow2) {
    for (int k = 2; k <= npow2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < npow2; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long x = d[i], y = d[ixj];
                    const bool up = ((i & k) == 0);
                    if ((x > y) == up) { d[i] = y; d[ixj] = x; }
                }
            }
            __syncthreads();
        }
    }
}

// The same network with one thread per compare-exchange PAIR, run by the first `nthr` threads of the block only
// (a multiple of 32; they meet at named barrier 1, the other warps go straight to the block barrier behind the
// sort): for the few hundred candidates of a typical image most of a 1024-thread block would only add barrier cost.
__device__ void bitonic_sort_pairs(unsigned long long* d, int npow2, int nthr) {
    const int half = npow2 >> 1;
    for (int k = 2; k <= npow2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < half; t += nthr) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int ixj = i | j;
                const unsigned long long x = d[i], y = d[ixj];
                const bool up = ((i & k) == 0);
                if ((x > y) == up) { d[i] = y; d[ixj] = x; }
            }
            asm volatile("bar.sync 1, %0;" ::"r"(nthr) : "memory");
        }
    }
}

// Radix selection (block-wide, every thread of the block calls it): among the first n keys of `keys`, find a 32-bit
// prefix T (the complemented score bits, i.e. the high word of a key) such that at least `want` keys have a prefix
// <= T and at most ~128 more than that do, unless ties make that impossible (then the full 32 bits are decided and
// all keys tied with T are included).  Returns T and the number of keys with prefix <= T.
// The bits all keys share (sign, exponent and leading mantissa bits of scores in (conf, 1]) are skipped: the block
// first reduces min and max of the prefixes, then histograms 11 bits at a time from the first bit that differs, so a
// few thousand keys spread over 2048 bins instead of piling onto a dozen.  The crossing bin is found with a block scan.
#define NMS_RADIX_BITS 11
__device__ unsigned nms_radix_threshold(const unsigned long long* keys, int n, int want, int* s_hist, int* s_out,
                                        int& n_le) {
    __shared__ int s_wt[32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nwarp = (blockDim.x + 31) >> 5;
    if (tid == 0) { s_out[0] = -1; s_out[1] = 0; }   // (as unsigned: min starts at 0xFFFFFFFF, max at 0)
    __syncthreads();
    {
        unsigned lo = 0xFFFFFFFFu, hi = 0u;
        for (int i = tid; i < n; i += blockDim.x) {
            const unsigned p = (unsigned)(keys[i] >> 32);
            lo = min(lo, p);
            hi = max(hi, p);
        }
        lo = __reduce_min_sync(0xffffffffu, lo);
        hi = __reduce_max_sync(0xffffffffu, hi);
        if (lane == 0) {
            atomicMin(reinterpret_cast<unsigned*>(&s_out[0]), lo);
            atomicMax(reinterpret_cast<unsigned*>(&s_out[1]), hi);
        }
    }
    __syncthreads();
    const unsigned kmin = (unsigned)s_out[0], kmax = (unsigned)s_out[1];
    __syncthreads();
    if (kmin == kmax) { n_le = n; return kmax; }
    int decided = __clz((int)(kmin ^ kmax));                  // leading bits common to all keys
    unsigned prefix = decided ? (kmax & ~(0xFFFFFFFFu >> decided)) : 0u;
    int before = 0;                                           // keys whose decided bits are smaller
    n_le = n;
    while (decided < 32) {
        const int w = min(NMS_RADIX_BITS, 32 - decided), nb = 1 << w, shift = 32 - decided - w;
        for (int i = tid; i < nb; i += blockDim.x) s_hist[i] = 0;
        __syncthreads();
        for (int i = tid; i < n; i += blockDim.x) {
            const unsigned p = (unsigned)(keys[i] >> 32);
            const bool in = (decided == 0) || ((p >> (32 - decided)) == (prefix >> (32 - decided)));
            if (in) atomicAdd(&s_hist[(p >> shift) & (nb - 1)], 1);
        }
        __syncthreads();
        // block scan over the bins: thread t owns bins [t*per, (t+1)*per)
        const int per = (nb + blockDim.x - 1) / blockDim.x;
        const int k0 = tid * per;
        int local = 0;
        for (int k = k0; k < min(k0 + per, nb); ++k) local += s_hist[k];
        int incl = local;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) s_wt[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            int v = (lane < nwarp) ? s_wt[lane] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, v, o);
                if (lane >= o) v += u;
            }
            s_wt[lane] = v;    // inclusive totals of the warps
        }
        __syncthreads();
        int cum = before + (wid ? s_wt[wid - 1] : 0) + incl - local;   // keys before this thread's first bin
        for (int k = k0; k < min(k0 + per, nb); ++k) {
            const int h = s_hist[k];
            if (cum < want && cum + h >= want) { s_out[0] = k; s_out[1] = cum; s_out[2] = cum + h; }
            cum += h;
        }
        __syncthreads();
        prefix |= (unsigned)s_out[0] << shift;
        before = s_out[1];
        n_le = s_out[2];
        decided += w;
        __syncthreads();
        if (n_le - want <= 128) break;   // close enough: take the whole selected bin
    }
    // undecided low bits: everything inside the selected bin counts as <= T
    return (decided < 32) ? (prefix | (0xFFFFFFFFu >> decided)) : prefix;
}

// per image: sort candidates by (score desc, input order asc) = stable descending sort
// (torchvision nms_kernel: scores.sort(0, descending=True)); cut to max_nms (utils/ops.py:401-402);
// emit class-offset boxes `x[:, :4] + cls * max_wh` in fp32 (utils/ops.py:405-406).
// When far more candidates pass the filter than max_nms keeps (the validator's conf 0.001 with multi-label:
// ~10^5..10^6 per image), the max_nms best are selected first (radix selection on the score bits, all keys
// tied with the threshold included) and only those are sorted.
//
// Greedy NMS walks the sorted list only until max_det boxes are kept, so with many candidates (more than NMS_TRANCHE)
// the first pass sorts just the leading tranche: the NMS_TRANCHE best are radix-selected (ties with the threshold
// included) and sorted in shared memory into ws.tkeys; ws.keys stays as the filter wrote it.  If the suppression runs
// out of tranche before max_det boxes are kept it raises ws.redo[b], and pass 1 of this kernel and of the suppression
// redo that image over all its candidates (the path every image took before) - same result either way.
__global__ void __launch_bounds__(1024) k_nms_sort(const float* __restrict__ pred, int CH, int A, ycr_nms_cfg_t cfg, NmsWs ws,
                                                   int fused_filter, const GatherFeats* __restrict__ gfp, int pass) {
    __shared__ unsigned long long s_keys[NMS_SORT_SMEM];
    __shared__ int s_out[4];
    const int b = blockIdx.x;
    if (pass == 0 && b == 0 && threadIdx.x == 0) *ws.done_ctr = 0;
    if (pass == 1 && !ws.redo[b]) return;
    unsigned long long* keys = ws.keys + (int64_t)b * ws.cap2;
    if (fused_filter && pass == 0) {
        // single-label NMS on a tensor whose best class per anchor came with the decode: the conf filter
        // (utils/ops.py:348, 386-387, 390-391) is one pass of this block over the image's 8-byte hints, so the
        // separate filter kernel, its global counter and its re-read of the keys are not needed
        __shared__ int s_cnt;
        if (threadIdx.x == 0) s_cnt = 0;
        __syncthreads();
        const int2* hint = reinterpret_cast<const int2*>(cfg.best_class) + (int64_t)b * A;
        const int nc = cfg.nc;
        for (int an = threadIdx.x; an < A; an += blockDim.x) {
            const int2 hv = hint[an];
            float best = __int_as_float(hv.x);
            int bc = hv.y;
            if (bc < 0) {   // no hint for this anchor: scan its class rows (first maximum)
                const float* p = pred + ((int64_t)b * CH + 4) * A + an;
                best = -3.4e38f;
                bc = 0;
                for (int c = 0; c < nc; ++c) {
                    const float v = p[(int64_t)c * A];
                    if (v > best) { best = v; bc = c; }
                }
            }
            bool ok = best > cfg.conf_thres;
            if (ok && cfg.classes) {
                ok = false;
                for (int k = 0; k < cfg.n_classes; ++k) ok = ok || (cfg.classes[k] == bc);
            }
            if (ok) {
                const int slot = atomicAdd(&s_cnt, 1);
                if (slot < ws.cap) keys[slot] = ((unsigned long long)(~__float_as_uint(best)) << 32) | (unsigned)(an * nc + bc);
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) ws.count[b] = s_cnt;
        __syncthreads();
    }
    const int n = min((fused_filter && pass == 0) ? *(volatile int*)&ws.count[b] : ws.count[b], ws.cap);
    if (threadIdx.x == 0) { ws.nsorted[b] = min(n, ws.nsel_cap); ws.tmode[b] = 0; }
    if (n == 0) return;
    bool sorted = false;
    const unsigned long long* skeys = keys;      // where the sorted list ends up
    int nsel = min(n, ws.nsel_cap);
    if (pass == 0 && ws.tkeys && n > NMS_TRANCHE && nsel > NMS_TRANCHE) {
        int n_le = 0;
        // (the selection returns up to 128 keys more than asked for: ask for that many fewer, so the tranche pads to 1024)
        const unsigned T = nms_radix_threshold(keys, n, NMS_TRANCHE - 128, reinterpret_cast<int*>(s_keys), s_out, n_le);
        if (n_le <= NMS_SORT_SMEM && n_le < n) {   // (block-uniform) else: too many ties, or nothing behind the tranche
            if (threadIdx.x == 0) s_out[3] = 0;
            __syncthreads();
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                const unsigned long long k = keys[i];
                if ((unsigned)(k >> 32) <= T) s_keys[atomicAdd(&s_out[3], 1)] = k;
            }
            __syncthreads();
            const int m = s_out[3];     // == n_le
            int mp2 = 1;
            while (mp2 < m) mp2 <<= 1;
            for (int i = m + threadIdx.x; i < mp2; i += blockDim.x) s_keys[i] = 0xFFFFFFFFFFFFFFFFull;
            __syncthreads();
            const int nthr = min((int)blockDim.x, max(32, mp2 >> 1));
            if ((int)threadIdx.x < nthr) bitonic_sort_pairs(s_keys, mp2, nthr);
            __syncthreads();
            unsigned long long* tk = ws.tkeys + (int64_t)b * NMS_SORT_SMEM;
            for (int i = threadIdx.x; i < m; i += blockDim.x) tk[i] = s_keys[i];
            nsel = min(m, ws.nsel_cap);
            if (threadIdx.x == 0) { ws.nsorted[b] = nsel; ws.tmode[b] = 1; }
            skeys = tk;
            sorted = true;
            __syncthreads();
        }
    }
    if (!sorted && ws.keys2 && n > NMS_PRESEL_MIN && n > ws.nsel_cap) {
        int n_le = 0;
        const unsigned T = nms_radix_threshold(keys, n, ws.nsel_cap, reinterpret_cast<int*>(s_keys), s_out, n_le);
        if (n_le <= ws.presel_cap2) {   // (block-uniform) otherwise too many ties: sort everything
            unsigned long long* sel = ws.keys2 + (int64_t)b * ws.presel_cap2;
            if (threadIdx.x == 0) s_out[3] = 0;
            __syncthreads();
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                const unsigned long long k = keys[i];
                if ((unsigned)(k >> 32) <= T) sel[atomicAdd(&s_out[3], 1)] = k;
            }
            __syncthreads();
            const int m = s_out[3];     // == n_le
            int mp2 = 1;
            while (mp2 < m) mp2 <<= 1;
            for (int i = m + threadIdx.x; i < mp2; i += blockDim.x) sel[i] = 0xFFFFFFFFFFFFFFFFull;
            __syncthreads();
            bitonic_sort(sel, mp2);
            for (int i = threadIdx.x; i < ws.nsel_cap; i += blockDim.x) keys[i] = sel[i];
            __syncthreads();
            sorted = true;
        }
    }
    if (!sorted) {
        int np2 = 1;
        while (np2 < n) np2 <<= 1;
        if (np2 <= NMS_SORT_SMEM) {
            for (int i = threadIdx.x; i < np2; i += blockDim.x) s_keys[i] = (i < n) ? keys[i] : 0xFFFFFFFFFFFFFFFFull;
            __syncthreads();
            const int nthr = min((int)blockDim.x, max(32, np2 >> 1));   // (block-uniform: n is)
            if ((int)threadIdx.x < nthr) bitonic_sort_pairs(s_keys, np2, nthr);
            __syncthreads();
            for (int i = threadIdx.x; i < n; i += blockDim.x) keys[i] = s_keys[i];
        } else {
            for (int i = n + threadIdx.x; i < np2; i += blockDim.x) keys[i] = 0xFFFFFFFFFFFFFFFFull;
            __syncthreads();
            bitonic_sort(keys, np2);
        }
    }
    __syncthreads();
    const int nc = cfg.nc;
    float4* boxes = ws.boxes + (int64_t)b * ws.nsel_cap;
    const float* p = pred ? pred + (int64_t)b * CH * A : nullptr;
    for (int i = threadIdx.x; i < nsel; i += blockDim.x) {
        const unsigned long long k = skeys[i];
        if (k == 0xFFFFFFFFFFFFFFFFull) { boxes[i] = make_float4(0.f, 0.f, -1.f, -1.f); continue; }
        const unsigned idx = (unsigned)(k & 0xFFFFFFFFull);
        const int an = idx / nc, c = idx - an * nc;
        const float off = cfg.agnostic ? 0.f : __fmul_rn((float)c, cfg.max_wh);
        // (the detect path has no prediction tensor: the box comes from the anchor's rays, same arithmetic)
        const float4 bx = p ? make_float4(p[an], p[(int64_t)A + an], p[(int64_t)2 * A + an], p[(int64_t)3 * A + an])
                            : box_from_feats(*gfp, b, an, nc);
        boxes[i] = make_float4(__fadd_rn(bx.x, off), __fadd_rn(bx.y, off), __fadd_rn(bx.z, off), __fadd_rn(bx.w, off));
    }
}

__device__ __forceinline__ bool iou_gt(const float4 a, const float area_a, const float4 b, const float thr) {
    // torchvision nms: inter / (area_a + area_b - inter) > thr, fp32, no epsilon
    const float w = fmaxf(0.f, fminf(a.z, b.z) - fmaxf(a.x, b.x));
    const float h = fmaxf(0.f, fminf(a.w, b.w) - fmaxf(a.y, b.y));
    const float inter = __fmul_rn(w, h);
    // disjoint boxes (most pairs): 0 / union is 0 or NaN, never above a threshold >= 0 - and a zero numerator sends
    // the IEEE division to its slow path
    if (!(inter > 0.f)) return false;
    const float area_b = __fmul_rn(b.z - b.x, b.w - b.y);
    const float iou = __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter));
    return iou > thr;
}

#define NMS_NT 256
#define NMS_MAX_KEEP 1024

// Called by warp 0 of every block of the FINAL suppression launch when the block is done (its out_counts entry is
// written): the last block to arrive turns the kept counts into the first output row of every image, so the gather
// blocks need not each add up the counts of the images before theirs.
__device__ __forceinline__ void nms_block_done(const NmsWs& ws, const int* out_counts, int B) {
    const int lane = threadIdx.x & 31;
    int last = 0;
    if (lane == 0) {
        __threadfence();
        last = (atomicAdd(ws.done_ctr, 1) == B - 1) ? 1 : 0;
    }
    last = __shfl_sync(0xffffffffu, last, 0);
    if (!last) return;
    __threadfence();
    int carry = 0;
    for (int i0 = 0; i0 < B; i0 += 32) {
        const int i = i0 + lane;
        const int c = (i < B) ? *(volatile const int*)&out_counts[i] : 0;
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (i < B) ws.row_base[i] = carry + incl - c;
        carry += __shfl_sync(0xffffffffu, incl, 31);
    }
}

// per image greedy suppression over the sorted candidates, 64 at a time ("pull" form of
// torchvision's greedy loop: a box is kept iff no EARLIER KEPT box overlaps it by more than thr):
//   (P) the 64 boxes of the chunk are tested against every box kept so far (4 threads per box),
//   (A) 64x64 bitmask inside the chunk,
//   (B) one thread walks the chunk in score order with bit operations and appends the survivors.
// The loop ends when max_det boxes are kept (utils/ops.py:408), so boxes after the max_det-th
// survivor are never touched.  Then the kept rows are gathered:
// [box xyxy | conf | class | nm mask channels] (utils/ops.py:383-387, 418).
__global__ void __launch_bounds__(NMS_NT) k_nms_suppress(const float* __restrict__ pred, int CH, int A, ycr_nms_cfg_t cfg, NmsWs ws,
                                                         float* __restrict__ out_rows, int* __restrict__ out_counts, int pass,
                                                         int final_launch) {
    __shared__ float4 s_kbox[NMS_MAX_KEEP];
    __shared__ int s_kept[NMS_MAX_KEEP];
    __shared__ unsigned long long s_mask[64];
    __shared__ float4 s_box[64];
    __shared__ unsigned long long s_dead;
    __shared__ int s_nkept, s_cut;
    const int b = blockIdx.x, tid = threadIdx.x;
    if (pass == 1 && !ws.redo[b]) {
        if (final_launch && tid < 32) nms_block_done(ws, out_counts, gridDim.x);
        return;
    }
    const int n_all = min(min(ws.count[b], ws.cap), ws.nsel_cap);
    const int tmode = ws.tmode[b];
    const int n = tmode ? min(ws.nsorted[b], n_all) : n_all;   // the sorted entries at hand (the leading tranche, or all)
    const int max_det = min(cfg.max_det, NMS_MAX_KEEP);
    const float4* boxes = ws.boxes + (int64_t)b * ws.nsel_cap;
    const unsigned long long* keys = tmode ? ws.tkeys + (int64_t)b * NMS_SORT_SMEM : ws.keys + (int64_t)b * ws.cap2;
    if (tid == 0) { s_nkept = 0; s_cut = n; }
    __syncthreads();
    if (cfg.classes) {  // entries of filtered-out classes were sorted to the end: drop them
        for (int i = tid; i < n; i += NMS_NT)
            if (keys[i] == 0xFFFFFFFFFFFFFFFFull) atomicMin(&s_cut, i);
        __syncthreads();
    }
    const int n_eff = s_cut;
    const float thr = cfg.iou_thres;
    for (int c0 = 0; c0 < n_eff; c0 += 64) {
        const int cn = min(64, n_eff - c0);
        if (tid < 64) {
            s_mask[tid] = 0ull;
            if (tid < cn) s_box[tid] = boxes[c0 + tid];
        }
        if (tid == 0) s_dead = 0ull;
        __syncthreads();
        const int nk = s_nkept;
        const int i = tid >> 2, q = tid & 3;
        if (i < cn) {
            const float4 bi = s_box[i];
            const float ai = __fmul_rn(bi.z - bi.x, bi.w - bi.y);
            // (P) against the boxes kept so far; kept box is the first IoU operand as in the reference loop
            bool dead = false;
            for (int k = q; k < nk && !dead; k += 4) {
                const float4 bk = s_kbox[k];
                const float ak = __fmul_rn(bk.z - bk.x, bk.w - bk.y);
                const float w = fmaxf(0.f, fminf(bk.z, bi.z) - fmaxf(bk.x, bi.x));
                const float h = fmaxf(0.f, fminf(bk.w, bi.w) - fmaxf(bk.y, bi.y));
                const float inter = __fmul_rn(w, h);
                dead = (inter > 0.f) && (__fdiv_rn(inter, __fsub_rn(__fadd_rn(ak, ai), inter)) > thr);
            }
            if (dead) atomicOr(&s_dead, 1ull << i);
            // (A) inside the chunk: row i of the (symmetric) overlap matrix, 16 columns per thread
            unsigned long long m = 0ull;
            for (int j = q * 16; j < q * 16 + 16; ++j)
                if (j != i && j < cn && iou_gt(bi, ai, s_box[j], thr)) m |= (1ull << j);
            if (m) atomicOr(&s_mask[i], m);
        }
        __syncthreads();
        if (tid < 32) {
            // (B) one warp decides the chunk, two boxes per lane (j = lane, lane + 32).  A box is removed once an
            // EARLIER box of the chunk that overlaps it is kept, and kept once all its earlier overlappers are
            // removed - the greedy order, resolved in rounds (the lowest undecided box is decided in every round; the
            // usual chunk needs two or three) instead of one survivor at a time.
            const unsigned long long dead0 = s_dead | ((cn < 64) ? (~0ull << cn) : 0ull);
            const unsigned long long e0 = s_mask[tid] & ((1ull << tid) - 1ull);                  // earlier overlappers of box tid
            const unsigned long long e1 = s_mask[tid + 32] & ((1ull << (tid + 32)) - 1ull);      // ... of box tid + 32
            unsigned long long kept = 0ull, rem = dead0;
            for (;;) {
                const unsigned long long und = ~(kept | rem);
                if (und == 0ull) break;
                bool k0 = false, r0 = false, k1 = false, r1 = false;
                if ((und >> tid) & 1ull) { r0 = (e0 & kept) != 0ull; k0 = !r0 && (e0 & ~rem) == 0ull; }
                if ((und >> (tid + 32)) & 1ull) { r1 = (e1 & kept) != 0ull; k1 = !r1 && (e1 & ~rem) == 0ull; }
                kept |= (unsigned long long)__ballot_sync(0xffffffffu, k0) | ((unsigned long long)__ballot_sync(0xffffffffu, k1) << 32);
                rem |= (unsigned long long)__ballot_sync(0xffffffffu, r0) | ((unsigned long long)__ballot_sync(0xffffffffu, r1) << 32);
            }
            // survivors in score order behind the nk boxes kept so far, cut at max_det (utils/ops.py:408)
            const int p0 = nk + __popcll(kept & ((1ull << tid) - 1ull));
            const int p1 = nk + __popcll(kept & ((1ull << (tid + 32)) - 1ull));
            if (((kept >> tid) & 1ull) && p0 < max_det) { s_kbox[p0] = s_box[tid]; s_kept[p0] = c0 + tid; }
            if (((kept >> (tid + 32)) & 1ull) && p1 < max_det) { s_kbox[p1] = s_box[tid + 32]; s_kept[p1] = c0 + tid + 32; }
            if (tid == 0) s_nkept = min(nk + __popcll(kept), max_det);
        }
        __syncthreads();
        if (s_nkept >= max_det) break;
    }
    __syncthreads();
    const int nk = s_nkept;
    if (tid == 0) {
        out_counts[b] = nk;
        // the tranche ended before max_det boxes were kept and candidates remain behind it: this image is redone
        if (pass == 0) {
            const int more = ws.more ? ws.more[b] : 0;                       // entries the limited filter left out
            const int n_exist = min(min(ws.count[b], ws.cap) + more, ws.nsel_cap);
            const int again = (nk < max_det && n < n_exist) ? 1 : 0;
            ws.redo[b] = again;
            if (again && more > 0) ws.count[b] = 0;                          // the plain filter refills this image
        }
    }
    const int nc = cfg.nc;
    // hand the kept list to the gather kernel: sorted position -> (anchor, class, score)
    int4* ko = ws.kept + (int64_t)b * NMS_MAX_KEEP;
    for (int r = tid; r < nk; r += NMS_NT) {
        const unsigned long long k = keys[s_kept[r]];
        const unsigned idx = (unsigned)(k & 0xFFFFFFFFull);
        const int an = idx / nc;
        ko[r] = make_int4(an, (int)(idx - an * nc), (int)(~(unsigned)(k >> 32)), 0);
    }
    if (final_launch && tid < 32) nms_block_done(ws, out_counts, gridDim.x);   // (out_counts[b] was written by lane 0 above)
}

// kept rows [box xyxy | conf | class | nm mask channels] (utils/ops.py:383-387, 418): the input is
// channel-major, so every element of a row is its own 32-byte sector; one thread per element keeps as
// many of those scattered loads in flight as the machine allows.
__global__ void __launch_bounds__(256) k_nms_gather(const float* __restrict__ pred, int CH, int A, ycr_nms_cfg_t cfg, NmsWs ws,
                                                    const int* __restrict__ counts, float* __restrict__ out_rows) {
    const int b = blockIdx.y;
    const int nc = cfg.nc, W = CH - 4 - nc + 6;
    const int nk = counts[b];
    if ((int)blockIdx.x * 256 >= nk * W) return;   // block-uniform
    // compact layout: first output row of image b = rows kept by the images before it (nms_block_done)
    const int64_t first = cfg.compact_rows ? (int64_t)ws.row_base[b] : (int64_t)b * cfg.max_det;
    const int e = blockIdx.x * 256 + threadIdx.x;
    if (e >= nk * W) return;
    const int r = e / W, col = e - r * W;
    const int4 k = ws.kept[(int64_t)b * NMS_MAX_KEEP + r];
    const float* p = pred + (int64_t)b * CH * A + k.x;
    float v;
    if (col < 4) v = p[(int64_t)col * A];
    else if (col == 4) v = __int_as_float(k.z);
    else if (col == 5) v = (float)k.y;
    else v = p[(int64_t)(4 + nc + col - 6) * A];
    out_rows[(first + r) * W + col] = v;
}

// The kept rows recomputed from the head feature maps the prediction was decoded from (when the caller still has
// them): one warp per kept row, lanes over the rays - R strided loads instead of 4 + 3R, identical arithmetic to
// k_decode*, the row written as one contiguous run.
__global__ void __launch_bounds__(256) k_nms_gather_feats(const __grid_constant__ GatherFeats gf, int CH, ycr_nms_cfg_t cfg, NmsWs ws,
                                                          const int* __restrict__ counts, float* __restrict__ out_rows) {
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int nk = counts[b];
    if (blockIdx.x * 8 >= nk) return;   // block-uniform
    if (r >= nk) return;
    const int64_t first = cfg.compact_rows ? (int64_t)ws.row_base[b] : (int64_t)b * cfg.max_det;
    const int R = gf.R, nc = cfg.nc, W = CH - 4 - nc + 6;
    const int4 k = ws.kept[(int64_t)b * NMS_MAX_KEEP + r];
    const int an = k.x;
    int l = 0;
#pragma unroll
    for (int q = 1; q < YCR_MAX_LEVELS; ++q)
        if (q < gf.grid.n_levels && an >= gf.grid.off[q]) l = q;
    const int hw = gf.grid.h[l] * gf.grid.w[l];
    const int al = an - gf.grid.off[l];
    const int iy = al / gf.grid.w[l], ix = al - iy * gf.grid.w[l];
    const float stride = gf.grid.stride[l];
    const float ax = ((float)ix + 0.5f) * stride, ay = ((float)iy + 0.5f) * stride;
    const int64_t f0 = (int64_t)b * (R + nc) * hw + al;
    float* o = out_rows + (first + r) * W;
    float minx = 3.4e38f, miny = 3.4e38f, maxx = -3.4e38f, maxy = -3.4e38f;
    for (int i = lane; i < R; i += 32) {
        const float dist = fmaxf(__fmul_rn(ycr_ld(gf.feats[l], f0 + (int64_t)i * hw, gf.dtype), stride), YCR_FLOOR);
        const float x = __fadd_rn(__fmul_rn(dist, gf.cs[i]), ax);
        const float y = __fadd_rn(__fmul_rn(dist, gf.cs[R + i]), ay);
        minx = fminf(minx, x); maxx = fmaxf(maxx, x);
        miny = fminf(miny, y); maxy = fmaxf(maxy, y);
        o[6 + i] = x;
        o[6 + R + i] = y;
        o[6 + 2 * R + i] = (dist > 1.f) ? 1.f : 0.f;
    }
#pragma unroll
    for (int q = 16; q > 0; q >>= 1) {
        minx = fminf(minx, __shfl_xor_sync(0xffffffffu, minx, q));
        miny = fminf(miny, __shfl_xor_sync(0xffffffffu, miny, q));
        maxx = fmaxf(maxx, __shfl_xor_sync(0xffffffffu, maxx, q));
        maxy = fmaxf(maxy, __shfl_xor_sync(0xffffffffu, maxy, q));
    }
    if (lane == 0) {
        o[0] = minx; o[1] = miny; o[2] = maxx; o[3] = maxy;
        o[4] = __int_as_float(k.z);
        o[5] = (float)k.y;
    }
}

// Deployment path (ycr_detect): best class per anchor straight from the class logits - no prediction tensor.
template <typename T>
__global__ void __launch_bounds__(256, DCB_MINB) k_cls_best_v4(const __grid_constant__ DecodeArgs d, int2* __restrict__ best) {
    const int A = d.grid.off[YCR_MAX_LEVELS];
    const int b = blockIdx.y;
    const int an = (blockIdx.x * 256 + threadIdx.x) * 4;
    if (an >= A) return;
    int l = 0;
#pragma unroll
    for (int k = 1; k < YCR_MAX_LEVELS; ++k)
        if (k < d.grid.n_levels && an >= d.grid.off[k]) l = k;
    const int hw = d.grid.h[l] * d.grid.w[l];
    const int al = an - d.grid.off[l];
    const int R = d.R, nc = d.nc;
    const T* fc = reinterpret_cast<const T*>(d.feats[l]) + (int64_t)b * (R + nc) * hw + (int64_t)R * hw + al;
    float bs[4] = {-3.4e38f, -3.4e38f, -3.4e38f, -3.4e38f};
    int bc[4] = {0, 0, 0, 0};
#pragma unroll 2
    for (int c = 0; c < nc; ++c) {
        const float4 x = YcrType<T>::ld4cs(fc + (int64_t)c * hw);
        float4 p;   // the sigmoid of k_decode_cls_best_v4, so that scores (and ties between classes) are the same
        p.x = 1.f / (1.f + expf(-x.x)); p.y = 1.f / (1.f + expf(-x.y));
        p.z = 1.f / (1.f + expf(-x.z)); p.w = 1.f / (1.f + expf(-x.w));
        if (p.x > bs[0]) { bs[0] = p.x; bc[0] = c; }
        if (p.y > bs[1]) { bs[1] = p.y; bc[1] = c; }
        if (p.z > bs[2]) { bs[2] = p.z; bc[2] = c; }
        if (p.w > bs[3]) { bs[3] = p.w; bc[3] = c; }
    }
    int2* bo = best + (int64_t)b * A + an;
    reinterpret_cast<int4*>(bo)[0] = make_int4(__float_as_int(bs[0]), bc[0], __float_as_int(bs[1]), bc[1]);
    reinterpret_cast<int4*>(bo)[1] = make_int4(__float_as_int(bs[2]), bc[2], __float_as_int(bs[3]), bc[3]);
}

static void fill_gather_feats(GatherFeats& gf, const ycr_grid_t* grid, const void* const* feats, int dtype, int R) {
    gf.grid = make_grid_dev(grid);
    for (int l = 0; l < grid->n_levels; ++l) gf.feats[l] = feats[l];
    gf.dtype = dtype;
    gf.R = R;
    for (int i = 0; i < R; ++i) {   // as launch_decode
        const float deg = (float)(i * (360 / R));
        const float ang = (deg / 180.f) * (float)3.141592653589793;
        gf.cs[i] = (float)cos((double)ang);
        gf.cs[R + i] = (float)sin((double)ang);
    }
}

size_t detect_workspace_bytes(int B, int A, const ycr_nms_cfg_t* cfg) {
    return nms_ws_layout(nullptr, nullptr, B, A, cfg) + align_up((size_t)B * A * sizeof(int2), 256) + align_up(sizeof(GatherFeats), 256);
}

// feats -> kept rows without ever writing the (B, 4+nc+3R, A) prediction: class pass, filter + sort (boxes from the
// rays of the candidates only), suppression, rows from the rays of the kept anchors.  Single-label (predictor) form.
int launch_detect(const ycr_grid_t* grid, const void* const* feats, int dtype, int B, int nc, int R, const ycr_nms_cfg_t* cfg,
                  float* out_rows, int* out_counts, void* workspace, size_t workspace_bytes, cudaStream_t st) {
    if (cfg->max_det < 1 || cfg->max_det > NMS_MAX_KEEP) { ycr_set_error("max_det = %d: at most %d", cfg->max_det, NMS_MAX_KEEP); return YCR_E_ARG; }
    if (cfg->multi_label && nc > 1) { ycr_set_error("ycr_detect is the single-label (predictor) form; use ycr_decode + ycr_nms for multi_label"); return YCR_E_ARG; }
    GatherFeats gf{};
    fill_gather_feats(gf, grid, feats, dtype, R);
    const int A = gf.grid.off[YCR_MAX_LEVELS];
    for (int l = 0; l < grid->n_levels; ++l)
        if ((grid->h[l] * grid->w[l]) % 4 || reinterpret_cast<uintptr_t>(feats[l]) % (4 * ycr_dtype_size(dtype))) {
            ycr_set_error("ycr_detect needs level sizes H*W that are multiples of 4 and maps aligned to 4 elements");
            return YCR_E_ARG;
        }
    if (detect_workspace_bytes(B, A, cfg) > workspace_bytes) { ycr_set_error("detect workspace too small"); return YCR_E_WORKSPACE; }
    NmsWs ws;
    char* base = reinterpret_cast<char*>(workspace);
    const size_t off1 = nms_ws_layout(&ws, base, B, A, cfg);
    int2* best = reinterpret_cast<int2*>(base + off1);
    GatherFeats* gf_d = reinterpret_cast<GatherFeats*>(base + off1 + align_up((size_t)B * A * sizeof(int2), 256));
    YCR_CUDA_CHECK(cudaMemcpyAsync(gf_d, &gf, sizeof(gf), cudaMemcpyHostToDevice, st));
    DecodeArgs d{};
    d.grid = gf.grid;
    for (int l = 0; l < grid->n_levels; ++l) d.feats[l] = feats[l];
    d.B = B; d.nc = nc; d.R = R; d.dtype = dtype;
    ycr_nms_cfg_t c = *cfg;
    c.best_class = best;
    c.nc = nc;
    const int CH = 4 + nc + 3 * R;
    {
        YcrProfScope ps(YCR_T_DECODE, st);
        dim3 g((A / 4 + 255) / 256, B);
        if (dtype == YCR_F16) k_cls_best_v4<__half><<<g, 256, 0, st>>>(d, best);
        else if (dtype == YCR_BF16) k_cls_best_v4<__nv_bfloat16><<<g, 256, 0, st>>>(d, best);
        else k_cls_best_v4<float><<<g, 256, 0, st>>>(d, best);
    }
    YCR_LAUNCH_CHECK();
    { YcrProfScope ps(YCR_T_NMS_SORT, st); k_nms_sort<<<B, 1024, 0, st>>>(nullptr, CH, A, c, ws, 1, gf_d, 0); }
    YCR_LAUNCH_CHECK();
    {
        YcrProfScope ps(YCR_T_NMS_SUPPRESS, st);
        k_nms_suppress<<<B, NMS_NT, 0, st>>>(nullptr, CH, A, c, ws, out_rows, out_counts, 0, ws.tkeys ? 0 : 1);
        if (ws.tkeys) {   // images whose leading tranche did not yield max_det boxes (blocks of the others return at once)
            k_nms_sort<<<B, 1024, 0, st>>>(nullptr, CH, A, c, ws, 1, gf_d, 1);
            k_nms_suppress<<<B, NMS_NT, 0, st>>>(nullptr, CH, A, c, ws, out_rows, out_counts, 1, 1);
        }
        dim3 gg((c.max_det + 7) / 8, B);
        k_nms_gather_feats<<<gg, 256, 0, st>>>(gf, CH, c, ws, out_counts, out_rows);
    }
    YCR_LAUNCH_CHECK();
    return YCR_OK;
}

size_t nms_workspace_bytes(int B, int A, const ycr_nms_cfg_t* cfg) { return nms_ws_layout(nullptr, nullptr, B, A, cfg); }

int launch_nms(const float* prediction, int B, int CH, int A, const ycr_nms_cfg_t* cfg, float* out_rows, int* out_counts,
               void* workspace, size_t workspace_bytes, cudaStream_t st) {
    if (cfg->max_det < 1 || cfg->max_det > NMS_MAX_KEEP) {
        ycr_set_error("max_det = %d: the suppression kernel keeps at most %d boxes per image in shared memory", cfg->max_det,
                      NMS_MAX_KEEP);
        return YCR_E_ARG;
    }
    if (B <= 0) return YCR_OK;                               // empty batch: nothing to do (utils/ops.py:362)
    if (A <= 0) { YCR_CUDA_CHECK(cudaMemsetAsync(out_counts, 0, (size_t)B * sizeof(int), st)); return YCR_OK; }
    NmsWs ws;
    const size_t need = nms_ws_layout(&ws, workspace, B, A, cfg);
    if (need > workspace_bytes) { ycr_set_error("nms workspace too small: need %zu have %zu", need, workspace_bytes); return YCR_E_WORKSPACE; }
    const int fused = (cfg->best_class && !(cfg->multi_label && cfg->nc > 1)) ? 1 : 0;
    const bool presel = !fused && ws.hist != nullptr;   // multi-label over a large (anchor, class) space: count first
    if (!fused) {
        dim3 g((A + 255) / 256, B);
        YcrProfScope ps(YCR_T_NMS_FILTER, st);
        if (presel) {
            YCR_CUDA_CHECK(cudaMemsetAsync(ws.hist, 0, (size_t)B * NMS_HBINS * sizeof(int), st));
            k_nms_hist<<<g, 256, 0, st>>>(prediction, CH, A, *cfg, ws);
            k_nms_pick<<<B, 256, 0, st>>>(ws);
        } else {
            YCR_CUDA_CHECK(cudaMemsetAsync(ws.count, 0, (size_t)(B + 1) * sizeof(int), st));
        }
        k_nms_filter<<<g, 256, 0, st>>>(prediction, CH, A, *cfg, ws, presel ? ws.bsel : nullptr, 0);
        YCR_LAUNCH_CHECK();
    }
    { YcrProfScope ps(YCR_T_NMS_SORT, st); k_nms_sort<<<B, 1024, 0, st>>>(prediction, CH, A, *cfg, ws, fused, nullptr, 0); }
    YCR_LAUNCH_CHECK();
    {
        YcrProfScope ps(YCR_T_NMS_SUPPRESS, st);
        k_nms_suppress<<<B, NMS_NT, 0, st>>>(prediction, CH, A, *cfg, ws, out_rows, out_counts, 0, ws.tkeys ? 0 : 1);
        if (ws.tkeys) {   // images whose leading tranche did not yield max_det boxes (blocks of the others return at once)
            if (presel) {
                dim3 g((A + 255) / 256, B);
                k_nms_filter<<<g, 256, 0, st>>>(prediction, CH, A, *cfg, ws, nullptr, 1);
            }
            k_nms_sort<<<B, 1024, 0, st>>>(prediction, CH, A, *cfg, ws, fused, nullptr, 1);
            k_nms_suppress<<<B, NMS_NT, 0, st>>>(prediction, CH, A, *cfg, ws, out_rows, out_counts, 1, 1);
        }
        const int maxk = cfg->max_det < NMS_MAX_KEEP ? cfg->max_det : NMS_MAX_KEEP;
        const int W = CH - 4 - cfg->nc + 6;
        const bool from_feats = cfg->grid && cfg->feats[0] && cfg->rays > 0 && cfg->rays <= 72 && CH == 4 + cfg->nc + 3 * cfg->rays;
        if (from_feats) {
            GatherFeats gf{};
            fill_gather_feats(gf, cfg->grid, cfg->feats, cfg->feats_dtype, cfg->rays);
            dim3 gg((maxk + 7) / 8, B);
            k_nms_gather_feats<<<gg, 256, 0, st>>>(gf, CH, *cfg, ws, out_counts, out_rows);
        } else {
            dim3 gg((maxk * W + 255) / 256, B);
            k_nms_gather<<<gg, 256, 0, st>>>(prediction, CH, A, *cfg, ws, out_counts, out_rows);
        }
    }
    YCR_LAUNCH_CHECK();
    return YCR_OK;
}
