// Contour -> mask rasterisation and mask IoU for the validator: the step the reference has
// commented out in ops.process_mask (utils/ops.py:794-809: per detection the valid contour points, truncated to
// int32, cv2.fillPoly) and metrics.mask_iou (utils/metrics.py:133-155).
//
// fillPoly semantics (OpenCV drawing.cpp; the parity tests pin them against cv2.fillPoly itself): every polygon edge is drawn as an 8-connected line (clipLine, then Bresenham left to right), and
// the interior is filled by even-odd scan lines in 16.16 fixed point - an edge covers the scan lines y0 <= y < y1
// with x = x0 + (y - y0) * dx, dx the truncated quotient; between the sorted crossings of a pair the pixels
// ceil(xa) .. floor(xb) are set.
#include "common.cuh"

#define RAS_NT 256
#define RAS_MAXV 72   // vertices per polygon (= rays)

struct RasEdge { int y0, y1; long long x, dx; };

__device__ __forceinline__ bool ras_clip_line(int w, int h, long long& x1, long long& y1, long long& x2, long long& y2) {
    const long long right = w - 1, bottom = h - 1;
    int c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
    int c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
        long long a;
        if (c1 & 12) {
            a = c1 < 8 ? 0 : bottom;
            x1 += (long long)((double)(a - y1) * (double)(x2 - x1) / (double)(y2 - y1));
            y1 = a;
            c1 = (x1 < 0) + (x1 > right) * 2;
        }
        if (c2 & 12) {
            a = c2 < 8 ? 0 : bottom;
            x2 += (long long)((double)(a - y2) * (double)(x2 - x1) / (double)(y2 - y1));
            y2 = a;
            c2 = (x2 < 0) + (x2 > right) * 2;
        }
        if ((c1 & c2) == 0 && (c1 | c2) != 0) {
            if (c1) {
                a = c1 == 1 ? 0 : right;
                y1 += (long long)((double)(a - x1) * (double)(y2 - y1) / (double)(x2 - x1));
                x1 = a;
                c1 = 0;
            }
            if (c2) {
                a = c2 == 1 ? 0 : right;
                y2 += (long long)((double)(a - x2) * (double)(y2 - y1) / (double)(x2 - x1));
                x2 = a;
                c2 = 0;
            }
        }
    }
    return (c1 | c2) == 0;
}

// one block per detection
__global__ void __launch_bounds__(RAS_NT) k_rasterize(const float* __restrict__ rows, int64_t row_stride, int R, int H, int W,
                                                      uint8_t* __restrict__ masks) {
    __shared__ int s_vx[RAS_MAXV], s_vy[RAS_MAXV];
    __shared__ RasEdge s_e[RAS_MAXV];
    __shared__ int s_nv, s_ne, s_ymin, s_ymax;
    const int d = blockIdx.x, tid = threadIdx.x;
    const float* r = rows + (int64_t)d * row_stride + 6;
    uint8_t* m = masks + (int64_t)d * H * W;
    if (tid == 0) {
        int nv = 0;
        for (int i = 0; i < R; ++i)
            if (r[2 * R + i] != 0.f) { s_vx[nv] = (int)r[i]; s_vy[nv] = (int)r[R + i]; ++nv; }   // astype(int32): truncation
        int ne = 0, ymin = 0x7fffffff, ymax = -0x7fffffff;
        for (int i = 0; i < nv; ++i) {
            const int j = (i == 0) ? nv - 1 : i - 1;
            const int x0 = s_vx[j], y0 = s_vy[j], x1 = s_vx[i], y1 = s_vy[i];
            if (y0 == y1) continue;
            RasEdge e;
            if (y0 < y1) { e.y0 = y0; e.y1 = y1; e.x = (long long)x0 << 16; }
            else { e.y0 = y1; e.y1 = y0; e.x = (long long)x1 << 16; }
            e.dx = (((long long)x1 - x0) << 16) / ((long long)y1 - y0);   // C division: truncated
            s_e[ne++] = e;
            ymin = min(ymin, e.y0);
            ymax = max(ymax, e.y1);
        }
        s_nv = nv; s_ne = ne; s_ymin = ymin; s_ymax = ymax;
    }
    // clear the mask while thread 0 builds the edge table
    {
        const int64_t n16 = ((int64_t)H * W) / 16;
        uint4* m4 = reinterpret_cast<uint4*>(m);
        if ((reinterpret_cast<uintptr_t>(m) & 15) == 0) {
            for (int64_t i = tid; i < n16; i += RAS_NT) m4[i] = make_uint4(0, 0, 0, 0);
            for (int64_t i = n16 * 16 + tid; i < (int64_t)H * W; i += RAS_NT) m[i] = 0;
        } else {
            for (int64_t i = tid; i < (int64_t)H * W; i += RAS_NT) m[i] = 0;
        }
    }
    __syncthreads();
    const int nv = s_nv, ne = s_ne;
    if (nv == 0) return;
    // scan-line fill: one warp per scan line
    const int lane = tid & 31, wid = tid >> 5;
    const int ya = max(s_ymin, 0), yb = min(s_ymax, H);
    for (int y = ya + wid; y < yb; y += RAS_NT / 32) {
        // crossings of this scan line, sorted: every lane builds the (short) list itself
        long long xs[8];
        int nx = 0;
        bool over = false;
        for (int k = 0; k < ne; ++k) {
            const RasEdge e = s_e[k];
            if (e.y0 <= y && y < e.y1) {
                const long long x = e.x + (long long)(y - e.y0) * e.dx;
                if (nx < 8) {
                    int p = nx++;
                    while (p > 0 && xs[p - 1] > x) { xs[p] = xs[p - 1]; --p; }
                    xs[p] = x;
                } else over = true;
            }
        }
        if (over) {
            // more than 8 crossings (a heavily folded contour): lane 0 walks the pairs with a selection scan
            if (lane == 0) {
                long long prev = -(1ll << 62);
                int prev_cnt = 0;   // crossings equal to prev already consumed
                bool open = false;
                long long xa = 0;
                for (;;) {
                    long long best = (1ll << 62);
                    int cnt = 0;
                    for (int k = 0; k < ne; ++k) {
                        const RasEdge e = s_e[k];
                        if (e.y0 <= y && y < e.y1) {
                            const long long x = e.x + (long long)(y - e.y0) * e.dx;
                            if (x > prev) { if (x < best) { best = x; cnt = 1; } else if (x == best) ++cnt; }
                        }
                    }
                    (void)prev_cnt;
                    if (cnt == 0) break;
                    for (int c = 0; c < cnt; ++c) {
                        if (!open) { xa = best; open = true; }
                        else {
                            int x1 = (int)((xa + 65535) >> 16), x2 = (int)(best >> 16);
                            if (x1 < W && x2 >= 0) { x1 = max(x1, 0); x2 = min(x2, W - 1); for (int x = x1; x <= x2; ++x) m[(int64_t)y * W + x] = 1; }
                            open = false;
                        }
                    }
                    prev = best;
                }
            }
            continue;
        }
        for (int k = 0; k + 1 < nx; k += 2) {
            int x1 = (int)((xs[k] + 65535) >> 16), x2 = (int)(xs[k + 1] >> 16);
            if (x1 < W && x2 >= 0) {
                x1 = max(x1, 0);
                x2 = min(x2, W - 1);
                for (int x = x1 + lane; x <= x2; x += 32) m[(int64_t)y * W + x] = 1;
            }
        }
    }
    // boundary lines: one thread per edge of the vertex ring
    for (int i = tid; i < nv; i += RAS_NT) {
        const int j = (i == 0) ? nv - 1 : i - 1;
        long long x0 = s_vx[j], y0 = s_vy[j], x1 = s_vx[i], y1 = s_vy[i];
        if (!ras_clip_line(W, H, x0, y0, x1, y1)) continue;
        long long dx = x1 - x0, dy = y1 - y0;
        if (dx < 0) { x0 = x1; y0 = y1; dx = -dx; dy = -dy; }
        const int sy = (dy >= 0) ? 1 : -1;
        dy = (dy < 0) ? -dy : dy;
        const bool steep = dy > dx;
        const long long a = steep ? dy : dx, b = steep ? dx : dy;
        long long err = a - 2 * b;
        int x = (int)x0, y = (int)y0;
        for (long long s = 0; s <= a; ++s) {
            if (x >= 0 && x < W && y >= 0 && y < H) m[(int64_t)y * W + x] = 1;
            const bool mm = err < 0;
            err += -2 * b + (mm ? 2 * a : 0);
            if (steep) { y += sy; x += mm ? 1 : 0; }
            else { x += 1; y += mm ? sy : 0; }
        }
    }
}

int launch_rasterize(const float* rows, int64_t row_stride, int n, int R, int H, int W, uint8_t* masks, cudaStream_t st) {
    if (n == 0) return YCR_OK;
    k_rasterize<<<n, RAS_NT, 0, st>>>(rows, row_stride, R, H, W, masks);
    YCR_LAUNCH_CHECK();
    return YCR_OK;
}

// ---- mask IoU: bit-pack both mask sets, then popcount ---------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) k_pack_masks(const T* __restrict__ m, int64_t n, int64_t words, uint32_t* __restrict__ out,
                                                    int* __restrict__ area) {
    const int row = blockIdx.y;
    const T* src = m + (int64_t)row * n;
    uint32_t* dst = out + (int64_t)row * words;
    int cnt = 0;
    for (int64_t p0 = ((int64_t)blockIdx.x * 256 + (threadIdx.x & ~31)); p0 < n; p0 += (int64_t)gridDim.x * 256) {
        const int64_t p = p0 + (threadIdx.x & 31);
        const bool on = (p < n) && (src[p] != (T)0);
        const unsigned b = __ballot_sync(0xffffffffu, on);
        if ((threadIdx.x & 31) == 0) { dst[p0 >> 5] = b; cnt += __popc(b); }
    }
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&area[row], cnt);
}

__global__ void __launch_bounds__(128) k_mask_iou(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, const int* __restrict__ area_a,
                                                  const int* __restrict__ area_b, int64_t words, int M, float eps, float* __restrict__ iou) {
    __shared__ int s_red[4];
    const int i = blockIdx.y, j = blockIdx.x;
    const uint32_t* pa = a + (int64_t)i * words;
    const uint32_t* pb = b + (int64_t)j * words;
    int c = 0;
    for (int64_t w = threadIdx.x; w < words; w += 128) c += __popc(pa[w] & pb[w]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        const float inter = (float)(s_red[0] + s_red[1] + s_red[2] + s_red[3]);
        const float uni = ((float)area_a[i] + (float)area_b[j]) - inter;   // utils/metrics.py:154
        iou[(int64_t)i * M + j] = inter / (uni + eps);
    }
}

size_t mask_iou_workspace_bytes(int N, int M, int64_t n) {
    const int64_t words = (n + 31) / 32;
    return (size_t)(N + M) * words * 4 + (size_t)(N + M) * 4 + 512;
}

int launch_mask_iou(const void* m1, int dt1, const void* m2, int dt2, int N, int M, int64_t n, float eps, float* iou,
                    void* workspace, size_t workspace_bytes, cudaStream_t st) {
    if (N == 0 || M == 0) return YCR_OK;
    const int64_t words = (n + 31) / 32;
    if (mask_iou_workspace_bytes(N, M, n) > workspace_bytes) { ycr_set_error("mask_iou workspace too small"); return YCR_E_WORKSPACE; }
    uint32_t* pa = reinterpret_cast<uint32_t*>(workspace);
    uint32_t* pb = pa + (int64_t)N * words;
    int* area = reinterpret_cast<int*>(pb + (int64_t)M * words);
    YCR_CUDA_CHECK(cudaMemsetAsync(area, 0, (size_t)(N + M) * 4, st));
    const int64_t gx64 = (n + 255) / 256;
    const int gx = (int)(gx64 < 1024 ? gx64 : 1024);
    for (int s = 0; s < 2; ++s) {
        const void* m = s ? m2 : m1;
        const int dt = s ? dt2 : dt1, rowsn = s ? M : N;
        uint32_t* o = s ? pb : pa;
        int* ar = area + (s ? N : 0);
        dim3 g(gx, rowsn);
        if (dt == 0) k_pack_masks<uint8_t><<<g, 256, 0, st>>>(reinterpret_cast<const uint8_t*>(m), n, words, o, ar);
        else if (dt == 1) k_pack_masks<float><<<g, 256, 0, st>>>(reinterpret_cast<const float*>(m), n, words, o, ar);
        else { ycr_set_error("mask dtype must be uint8 (0) or float32 (1)"); return YCR_E_ARG; }
    }
    YCR_LAUNCH_CHECK();
    dim3 g2(M, N);
    k_mask_iou<<<g2, 128, 0, st>>>(pa, pb, area, area + N, words, M, eps, iou);
    YCR_LAUNCH_CHECK();
    return YCR_OK;
}
