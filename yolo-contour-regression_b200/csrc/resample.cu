// Contour resampling on the device: replaces ops.resample_segments (utils/ops.py:676-693, called from
// utils/instance.py:202 with n=360): close the polygon, then np.interp of x and y onto
// linspace(0, len, n).  Arithmetic restates numpy's in double, operation by operation (step = L/(n-1),
// x_k = k*step, last x = L; value = (f[j+1]-f[j])*(x-j) + f[j]), so results are bit-identical to the
// reference's float32 output.  SURVEY.md §8-f.3 ("next" row: on-wire GT format produced on the GPU).
#include "common.cuh"

__global__ void __launch_bounds__(128) k_resample(const float* __restrict__ pts, const int* __restrict__ offsets, int n_out,
                                                  float* __restrict__ out) {
    const int s = blockIdx.y;
    const int k = blockIdx.x * 128 + threadIdx.x;
    if (k >= n_out) return;
    const int o = offsets[s], m = offsets[s + 1] - o;  // m open-polygon points; the closed one has m+1
    float* dst = out + ((int64_t)s * n_out + k) * 2;
    if (m <= 0) { dst[0] = 0.f; dst[1] = 0.f; return; }
    const double L = (double)m;                          // len(closed) - 1
    const double step = __ddiv_rn(L, (double)(n_out - 1));
    const double x = (k == n_out - 1) ? L : __dmul_rn((double)k, step);
    int j = (int)x;                                      // xp = arange(m+1): interval index
    if (j >= m) {                                        // right edge: np.interp returns fp[-1] = first point
        dst[0] = pts[(int64_t)o * 2];
        dst[1] = pts[(int64_t)o * 2 + 1];
        return;
    }
    const int j1 = (j + 1 == m) ? 0 : j + 1;             // closing point wraps to the first
    const double t = __dsub_rn(x, (double)j);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        const double f0 = (double)pts[(int64_t)(o + j) * 2 + c], f1 = (double)pts[(int64_t)(o + j1) * 2 + c];
        const double slope = __ddiv_rn(__dsub_rn(f1, f0), 1.0);
        dst[c] = (float)__dadd_rn(__dmul_rn(slope, t), f0);
    }
}

int launch_resample(const float* pts, const int* offsets, int S, int n_out, float* out, cudaStream_t st) {
    if (S <= 0) return YCR_OK;
    dim3 grid((n_out + 127) / 128, S);
    k_resample<<<grid, 128, 0, st>>>(pts, offsets, n_out, out);
    YCR_LAUNCH_CHECK();
    return YCR_OK;
}
