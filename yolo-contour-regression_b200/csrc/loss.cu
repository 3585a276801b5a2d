// Training path, streaming stage: BCE-with-logits class loss (utils/loss.py:866-867) against the
// implicit sparse target_scores, the gradient of the whole loss with respect to every head output
// written in the same pass, the final deterministic reductions, and GT packing
// (utils/loss.py:215-239).  Paths under /root/reference/ultralytics-main/ultralytics/.
#include "train_path.cuh"

#define K5_NT 256

// One thread per anchor, looping over the channel dimension: every global access of a warp is a
// contiguous 128-byte line of the NCHW feature map (a_local is the fastest dimension).
__global__ void __launch_bounds__(K5_NT) k_loss_stream(const __grid_constant__ AssignArgs a, const __grid_constant__ AssignWs ws,
                                                       const void* f0, const void* f1, const void* f2, const void* f3,
                                                       void* g0, void* g1, void* g2, void* g3, float cls_gain) {
    pdl_enter();
    __shared__ float s_red[K5_NT / 32];
    const int A = a.grid.off[YCR_MAX_LEVELS];
    const int b = blockIdx.y;
    const int an = blockIdx.x * K5_NT + threadIdx.x;
    const int R = a.cfg.rays, nc = a.cfg.num_classes;
    float acc = 0.f;
    if (an < A) {
        int l = 0;
#pragma unroll
        for (int k = 1; k < YCR_MAX_LEVELS; ++k)
            if (k < a.grid.n_levels && an >= a.grid.off[k]) l = k;
        const void* f = (l == 0) ? f0 : (l == 1) ? f1 : (l == 2) ? f2 : f3;
        void* g = (l == 0) ? g0 : (l == 1) ? g1 : (l == 2) ? g2 : g3;
        const int dt = a.dtype;
        const int hw = a.grid.h[l] * a.grid.w[l];
        const int al = an - a.grid.off[l];
        const int64_t base = (int64_t)b * (R + nc) * hw + al;
        const int row = ws.pos_row[(int64_t)b * A + an];
        int label = -1;
        float tnorm = 0.f;
        const float* pg = nullptr;
        if (row >= 0) {
            const int gi = ws.pos_g[(int64_t)b * ws.pos_cap + row];
            int64_t lab = (int64_t)a.gt.labels[(int64_t)(b * a.gt.G + gi) * a.gt.labels_stride];
            label = (int)(lab < 0 ? 0 : lab);
            tnorm = ycr_round_to(ws.pos_norm[(int64_t)b * ws.pos_cap + row], dt);   // target_scores.to(dtype), utils/loss.py:867
            pg = ws.pos_grad + ((int64_t)b * ws.pos_cap + row) * R;
        }
        const float gscale = cls_gain * (float)a.gt.B / ws.tss[0];
        if (g) {
            for (int i = 0; i < R; ++i) ycr_st(g, base + (int64_t)i * hw, pg ? pg[i] : 0.f, dt);
        }
        const int64_t cbase = base + (int64_t)R * hw;
#pragma unroll 4
        for (int c = 0; c < nc; ++c) {
            const float x = ycr_ld(f, cbase + (int64_t)c * hw, dt);
            const float t = (c == label) ? tnorm : 0.f;
            const float e = __expf(-fabsf(x));
            acc += fmaxf(x, 0.f) - x * t + log1pf(e);
            if (g) {
                const float r = __fdividef(1.f, 1.f + e);
                const float sig = (x >= 0.f) ? r : e * r;
                ycr_st(g, cbase + (int64_t)c * hw, (sig - t) * gscale, dt);
            }
        }
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = (threadIdx.x < K5_NT / 32) ? s_red[threadIdx.x] : 0.f;
        v = warp_sum(v);
        if (threadIdx.x == 0) ws.bce_part[blockIdx.y * gridDim.x + blockIdx.x] = v;
    }
}

#ifndef K5V_NT
#define K5V_NT 256   // threads per block of the vectorised kernel (four anchors per thread)
#endif
#ifndef K5_MINB
#define K5_MINB 4   // 64 registers: four 256-thread blocks per SM
#endif
#define YCR_STR2(x) #x
#define YCR_PRAGMA_UNROLL(n) _Pragma(YCR_STR2(unroll n))
// Vectorised variant (every level's H*W is a multiple of 4, true for all image sizes divisible by 64):
// one thread per FOUR consecutive anchors, 128-bit loads and stores, four channels in flight.
__device__ __forceinline__ float bce_term(float x, float t, float gscale, float& grad) {
    const float e = __expf(-fabsf(x));
    const float r = __fdividef(1.f, 1.f + e);
    const float sig = (x >= 0.f) ? r : e * r;
    grad = (sig - t) * gscale;
    // log(1 + e), e in (0, 1]: the fast logarithm is within 4e-7 absolute there (terms average ~0.7 and the
    // loss is compared at 1e-5 relative); the accurate log1pf made this kernel instruction-bound
    return fmaxf(x, 0.f) - x * t + __logf(1.f + e);
}

template <typename T>
__global__ void __launch_bounds__(K5V_NT, K5_MINB) k_loss_stream_v4(const __grid_constant__ AssignArgs a, const __grid_constant__ AssignWs ws,
                                                          const T* f0, const T* f1, const T* f2, const T* f3,
                                                          T* g0, T* g1, T* g2, T* g3, float cls_gain) {
    pdl_enter();
    __shared__ float s_red[K5V_NT / 32];
    const int A = a.grid.off[YCR_MAX_LEVELS];
    const int b = blockIdx.y;
    const int an = (blockIdx.x * K5V_NT + threadIdx.x) * 4;
    const int R = a.cfg.rays, nc = a.cfg.num_classes;
    float acc = 0.f;
    if (an < A) {
        int l = 0;
#pragma unroll
        for (int k = 1; k < YCR_MAX_LEVELS; ++k)
            if (k < a.grid.n_levels && an >= a.grid.off[k]) l = k;
        const T* f = (l == 0) ? f0 : (l == 1) ? f1 : (l == 2) ? f2 : f3;
        T* g = (l == 0) ? g0 : (l == 1) ? g1 : (l == 2) ? g2 : g3;
        const int hw = a.grid.h[l] * a.grid.w[l];
        const int al = an - a.grid.off[l];
        const int64_t base = (int64_t)b * (R + nc) * hw + al;
        const int4 rows = *reinterpret_cast<const int4*>(ws.pos_row + (int64_t)b * A + an);
        const int row[4] = {rows.x, rows.y, rows.z, rows.w};
        int label[4] = {-1, -1, -1, -1};
        float tn[4] = {0.f, 0.f, 0.f, 0.f};
        const float* pg[4] = {nullptr, nullptr, nullptr, nullptr};
        const bool any_pos = (rows.x & rows.y & rows.z & rows.w) >= 0;  // some row index is non-negative
        if (any_pos) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (row[k] >= 0) {
                    const int gi = ws.pos_g[(int64_t)b * ws.pos_cap + row[k]];
                    const int64_t lab = (int64_t)a.gt.labels[(int64_t)(b * a.gt.G + gi) * a.gt.labels_stride];
                    label[k] = (int)(lab < 0 ? 0 : lab);
                    tn[k] = ycr_round_to(ws.pos_norm[(int64_t)b * ws.pos_cap + row[k]], YcrType<T>::code);   // target_scores.to(dtype)
                    pg[k] = ws.pos_grad + ((int64_t)b * ws.pos_cap + row[k]) * R;
                }
            }
        }
        const float gscale = cls_gain * (float)a.gt.B / ws.tss[0];
        if (g) {
            if (any_pos) {
                for (int i = 0; i < R; ++i) {
                    float4 v;
                    v.x = pg[0] ? pg[0][i] : 0.f; v.y = pg[1] ? pg[1][i] : 0.f;
                    v.z = pg[2] ? pg[2][i] : 0.f; v.w = pg[3] ? pg[3][i] : 0.f;
                    YcrType<T>::st4cs(g + base + (int64_t)i * hw, v);
                }
            } else {
                const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
                for (int i = 0; i < R; ++i) YcrType<T>::st4cs(g + base + (int64_t)i * hw, z);
            }
        }
        const T* fc = f + base + (int64_t)R * hw;
        T* gc = g ? g + base + (int64_t)R * hw : nullptr;
#ifndef K5_UNROLL
#define K5_UNROLL 4
#endif
        // K5_UNROLL class rows are requested before the first of them is used: the gradient stores may alias the
        // inputs as far as the compiler knows, so a load written behind a store stays behind it - with the plain
        // load / compute / store loop every thread had ONE load in flight
        for (int c0 = 0; c0 < nc; c0 += K5_UNROLL) {
            float4 x[K5_UNROLL];
#pragma unroll
            for (int u = 0; u < K5_UNROLL; ++u)
                if (c0 + u < nc) x[u] = YcrType<T>::ld4cs(fc + (int64_t)(c0 + u) * hw);
#pragma unroll
            for (int u = 0; u < K5_UNROLL; ++u) {
                const int c = c0 + u;
                if (c < nc) {
                    float4 gr;
                    acc += bce_term(x[u].x, (c == label[0]) ? tn[0] : 0.f, gscale, gr.x);
                    acc += bce_term(x[u].y, (c == label[1]) ? tn[1] : 0.f, gscale, gr.y);
                    acc += bce_term(x[u].z, (c == label[2]) ? tn[2] : 0.f, gscale, gr.z);
                    acc += bce_term(x[u].w, (c == label[3]) ? tn[3] : 0.f, gscale, gr.w);
                    if (gc) YcrType<T>::st4cs(gc + (int64_t)c * hw, gr);
                }
            }
        }
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = (threadIdx.x < K5V_NT / 32) ? s_red[threadIdx.x] : 0.f;
        v = warp_sum(v);
        if (threadIdx.x == 0) ws.bce_part[blockIdx.y * gridDim.x + blockIdx.x] = v;
    }
}

// total = (box_gain*l_box + cls_gain*l_cls) * B  (utils/loss.py:874-878); fixed-order reductions:
// warp w sums images w, w+32, ... and partials w, w+32.., then one fixed tree over the 32 warps.
__global__ void __launch_bounds__(1024) k_loss_finalize(AssignWs ws, int B, int n_bce, float box_gain, float cls_gain,
                                                        float* loss_out) {
    pdl_enter();
    __shared__ double s_b[32], s_p[32];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    double sb = 0.0, sp = 0.0;
    for (int i = tid; i < n_bce; i += 1024) sb += (double)ws.bce_part[i];
    for (int b = warp; b < B; b += 32) {
        const int n = ws.npos[b];
        for (int r = lane; r < n; r += 32) sp += (double)ws.pos_loss[(int64_t)b * ws.pos_cap + r];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sb += __shfl_xor_sync(0xffffffffu, sb, o);
        sp += __shfl_xor_sync(0xffffffffu, sp, o);
    }
    if (lane == 0) { s_b[warp] = sb; s_p[warp] = sp; }
    __syncthreads();
    if (tid == 0) {
        double tb = 0.0, tp = 0.0;
        for (int w = 0; w < 32; ++w) { tb += s_b[w]; tp += s_p[w]; }
        const double tss = (double)ws.tss[0];
        const float l_box = (float)(tp / tss) * box_gain;
        const float l_cls = (float)(tb / tss) * cls_gain;
        loss_out[0] = (l_box + l_cls) * (float)B;
        loss_out[1] = l_box;
        loss_out[2] = l_cls;
        loss_out[3] = (float)tss;
        if (ws.err[0]) {  // candidate capacity exceeded: the assignment was skipped - fail loudly, not silently
            const float nan = __int_as_float(0x7fc00000);
            loss_out[0] = loss_out[1] = loss_out[2] = nan;
        }
    }
}

template <typename T>
static cudaError_t launch_stream_v4(const AssignArgs& a, const AssignWs& ws, const void* const* f, void* const* g, dim3 grid,
                                    float cls_gain, cudaStream_t st) {
    return ycr_launch(k_loss_stream_v4<T>, grid, dim3(K5V_NT), 0, st, a, ws, reinterpret_cast<const T*>(f[0]),
                      reinterpret_cast<const T*>(f[1]), reinterpret_cast<const T*>(f[2]), reinterpret_cast<const T*>(f[3]),
                      reinterpret_cast<T*>(g[0]), reinterpret_cast<T*>(g[1]), reinterpret_cast<T*>(g[2]), reinterpret_cast<T*>(g[3]),
                      cls_gain);
}

int launch_loss_stream(const AssignArgs& a, const AssignWs& ws, const void* const* feats, void* const* grads,
                       const ycr_loss_cfg_t& lcfg, float* loss_out, cudaStream_t st) {
    const int B = a.gt.B;
    const int A = a.grid.off[YCR_MAX_LEVELS];
    const void* f[4] = {nullptr, nullptr, nullptr, nullptr};
    void* g[4] = {nullptr, nullptr, nullptr, nullptr};
    const uintptr_t align = 4 * ycr_dtype_size(a.dtype);   // four elements per access
    bool vec = true;
    for (int l = 0; l < a.grid.n_levels; ++l) {
        f[l] = feats[l];
        g[l] = grads ? grads[l] : nullptr;
        vec = vec && ((a.grid.h[l] * a.grid.w[l]) % 4 == 0) && (reinterpret_cast<uintptr_t>(f[l]) % align == 0) &&
              (!g[l] || reinterpret_cast<uintptr_t>(g[l]) % align == 0);
    }
    int nblk;
    if (vec) {
        dim3 grid((A / 4 + K5V_NT - 1) / K5V_NT, B);
        nblk = (int)(grid.x * grid.y);
        YcrProfScope ps(YCR_T_STREAM, st);
        if (a.dtype == YCR_F16) YCR_CUDA_CHECK(launch_stream_v4<__half>(a, ws, f, g, grid, lcfg.cls_gain, st));
        else if (a.dtype == YCR_BF16) YCR_CUDA_CHECK(launch_stream_v4<__nv_bfloat16>(a, ws, f, g, grid, lcfg.cls_gain, st));
        else YCR_CUDA_CHECK(launch_stream_v4<float>(a, ws, f, g, grid, lcfg.cls_gain, st));
    } else {
        dim3 grid((A + K5_NT - 1) / K5_NT, B);
        nblk = (int)(grid.x * grid.y);
        YcrProfScope ps(YCR_T_STREAM, st);
        YCR_CUDA_CHECK(ycr_launch(k_loss_stream, grid, dim3(K5_NT), 0, st, a, ws, f[0], f[1], f[2], f[3], g[0], g[1], g[2], g[3], lcfg.cls_gain));
    }
    YCR_LAUNCH_CHECK();
    { YcrProfScope ps(YCR_T_FINAL, st); YCR_CUDA_CHECK(ycr_launch(k_loss_finalize, dim3(1), dim3(1024), 0, st, ws, B, nblk, lcfg.box_gain, lcfg.cls_gain, loss_out)); }
    YCR_LAUNCH_CHECK();
    return YCR_OK;
}

// ------------------------------------------------------------------------------------------------
// grad *= scale (device scalar); exits at once when the scalar is 1
// ------------------------------------------------------------------------------------------------
struct ScaleArgs { void* p[YCR_MAX_LEVELS]; int64_t n[YCR_MAX_LEVELS]; int n_levels; int dtype; };

__global__ void __launch_bounds__(256) k_scale(const ScaleArgs sa, const float* scale) {
    pdl_enter();
    const float s = *scale;
    if (s == 1.f) return;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int l = 0; l < sa.n_levels; ++l)
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < sa.n[l]; i += stride)
            ycr_st(sa.p[l], i, ycr_ld(sa.p[l], i, sa.dtype) * s, sa.dtype);
}

int launch_scale(void* const* p, const int64_t* n, int n_levels, int dtype, const float* scale, cudaStream_t st) {
    ScaleArgs sa{};
    sa.n_levels = 0;
    sa.dtype = dtype;
    for (int l = 0; l < n_levels && l < YCR_MAX_LEVELS; ++l)
        if (n[l] > 0) { sa.p[sa.n_levels] = p[l]; sa.n[sa.n_levels] = n[l]; ++sa.n_levels; }
    if (sa.n_levels == 0) return YCR_OK;
    YCR_CUDA_CHECK(ycr_launch(k_scale, dim3(YCR_NUM_SMS * 2), dim3(256), 0, st, sa, scale));
    YCR_LAUNCH_CHECK();
    return YCR_OK;
}

// ------------------------------------------------------------------------------------------------
// GT packing: rows -> (B,G,725) padded, in px (utils/loss.py:215-239, 834-844).  Slot of row n inside
// its image = number of earlier rows with the same image index (stable, as `targets[matches]`).
// The image indices are first gathered into a compact array (one strided read per row), so the rank of a row
// is a scan of contiguous ints, not of the 726-float rows themselves.
// ------------------------------------------------------------------------------------------------
__global__ void k_pack_index(const float* __restrict__ head, int64_t hs, int N, int* __restrict__ idx) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n < N) idx[n] = (int)head[(int64_t)n * hs];
}

// head: [image index, class, x, y, w, h] per row (hs floats apart); seg: 720 contour values per row (ss floats apart)
__global__ void __launch_bounds__(128) k_pack_targets(const float* __restrict__ head, int64_t hs, const float* __restrict__ seg, int64_t ss,
                                                      const int* __restrict__ idx, int N, int B, int G, float img_w, float img_h,
                                                      float* __restrict__ out) {
    const int n = blockIdx.x;
    const float* t = head + (int64_t)n * hs;
    const float* sg = seg + (int64_t)n * ss;
    const int b = idx[n];
    __shared__ int s_part[4];
    int cnt = 0;
    for (int k = threadIdx.x; k < n; k += 128) cnt += (idx[k] == b) ? 1 : 0;
    cnt = (int)warp_sum((float)cnt);   // (n < 2^24: exact)
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = cnt;
    __syncthreads();
    const int slot = s_part[0] + s_part[1] + s_part[2] + s_part[3];
    if (b < 0 || b >= B || slot >= G) return;
    float* row = out + ((int64_t)b * G + slot) * (5 + 2 * YCR_C);
    if (threadIdx.x == 0) {
        row[0] = t[1];
        const float cx = t[2] * img_w, cy = t[3] * img_h, w = t[4] * img_w, h = t[5] * img_h;
        row[1] = cx - w / 2; row[2] = cy - h / 2; row[3] = cx + w / 2; row[4] = cy + h / 2;
    }
    // the reference scales the first 360 contour values by width and the last 360 by height although
    // the data is x,y-interleaved (utils/loss.py:236-237); restated literally
    for (int k = threadIdx.x; k < 2 * YCR_C; k += 128) row[5 + k] = sg[k] * ((k < YCR_C) ? img_w : img_h);
}

// The same with the row of every (image, slot) known (row_of[b*G+g], -1 = padding; the host-side staging computes it
// while it copies the rows): one block per padded row, no memset, no rank scan.
__global__ void __launch_bounds__(128) k_pack_targets_mapped(const float* __restrict__ head, int64_t hs, const float* __restrict__ seg,
                                                             int64_t ss, const int* __restrict__ row_of, float img_w, float img_h,
                                                             float* __restrict__ out) {
    pdl_enter();
    const int bg = blockIdx.x;
    const int n = row_of[bg];
    float* row = out + (int64_t)bg * (5 + 2 * YCR_C);
    if (n < 0) {
        for (int k = threadIdx.x; k < 5 + 2 * YCR_C; k += 128) row[k] = 0.f;
        return;
    }
    const float* t = head + (int64_t)n * hs;
    const float* sg = seg + (int64_t)n * ss;
    if (threadIdx.x == 0) {
        row[0] = t[1];
        const float cx = t[2] * img_w, cy = t[3] * img_h, w = t[4] * img_w, h = t[5] * img_h;
        row[1] = cx - w / 2; row[2] = cy - h / 2; row[3] = cx + w / 2; row[4] = cy + h / 2;
    }
    for (int k = threadIdx.x; k < 2 * YCR_C; k += 128) row[5 + k] = sg[k] * ((k < YCR_C) ? img_w : img_h);
}

int launch_pack_targets_mapped(const float* head, int64_t hs, const float* seg, int64_t ss, const int* row_of, int B, int G,
                               float img_w, float img_h, float* out, cudaStream_t st) {
    if (B * G == 0) return YCR_OK;
    YCR_CUDA_CHECK(ycr_launch(k_pack_targets_mapped, dim3(B * G), dim3(128), 0, st, head, hs, seg, ss, row_of, img_w, img_h, out));
    YCR_LAUNCH_CHECK();
    return YCR_OK;
}

int launch_pack_targets(const float* head, int64_t hs, const float* seg, int64_t ss, int N, int B, int G, float img_w, float img_h,
                        float* out, cudaStream_t st) {
    // the padded tensor must be zero where no row lands; the compact index array comes from the stream-ordered pool
    YCR_CUDA_CHECK(cudaMemsetAsync(out, 0, (size_t)B * G * (5 + 2 * YCR_C) * sizeof(float), st));
    if (N > 0) {
        int* idx = nullptr;
        YCR_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void**>(&idx), (size_t)N * sizeof(int), st));
        k_pack_index<<<(N + 255) / 256, 256, 0, st>>>(head, hs, N, idx);
        k_pack_targets<<<N, 128, 0, st>>>(head, hs, seg, ss, idx, N, B, G, img_w, img_h, out);
        YCR_LAUNCH_CHECK();
        YCR_CUDA_CHECK(cudaFreeAsync(idx, st));
    }
    return YCR_OK;
}
