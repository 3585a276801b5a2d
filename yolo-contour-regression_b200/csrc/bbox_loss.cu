// Box terms of the v8 loss family: CIoU + Distribution-Focal loss, forward and backward in one pass.
// Replaces BboxLoss.forward / _df_loss (utils/loss.py:53-87), bbox_iou(CIoU=True) (utils/metrics.py:77-130)
// and bbox2dist (utils/tal.py:1437-1440); paths under /root/reference/ultralytics-main/ultralytics/.
// The live polar v8SegmentationLoss constructs this loss (utils/loss.py:211) but never calls it; it is kept
// because the north_star names it and the `ori*` experiment classes of the fork use it.
#include "common.cuh"

#define BL_NT 128

struct BboxLossArgs {
    const float* pred_dist;      // (B,A,4*(reg_max+1)) logits
    const float* pred_bboxes;    // (B,A,4) xyxy
    const float* anchor_points;  // (A,2)
    const float* target_bboxes;  // (B,A,4) xyxy
    const float* target_scores;  // (B,A,nc)
    const uint8_t* fg_mask;      // (B,A)
    const float* tss;            // device scalar: target_scores_sum
    float* grad_dist;            // (B,A,4*(reg_max+1)) or null
    float* grad_bboxes;          // (B,A,4) or null
    float* part;                 // [2][nblocks] block partial sums (iou, dfl)
    int B, A, nc, reg_max, use_dfl;
};

__global__ void __launch_bounds__(BL_NT) k_bbox_loss(const BboxLossArgs p) {
    __shared__ float s_red[2][BL_NT / 32];
    const int64_t n = (int64_t)p.B * p.A;
    const int64_t i = (int64_t)blockIdx.x * BL_NT + threadIdx.x;
    const int bins = p.reg_max + 1;
    float l_iou = 0.f, l_dfl = 0.f;
    if (i < n) {
        const bool fg = p.fg_mask[i] != 0;
        float gb[4] = {0.f, 0.f, 0.f, 0.f};
        float* gd = p.grad_dist ? p.grad_dist + i * 4 * bins : nullptr;
        if (fg) {
            const float inv_tss = 1.f / p.tss[0];
            float w = 0.f;
            const float* ts = p.target_scores + i * p.nc;
            for (int c = 0; c < p.nc; ++c) w += ts[c];
            const float4 pb = reinterpret_cast<const float4*>(p.pred_bboxes)[i];
            const float4 tb = reinterpret_cast<const float4*>(p.target_bboxes)[i];
            const float eps = 1e-7f;
            // ---- CIoU (utils/metrics.py:104-128) ----
            const float w1 = pb.z - pb.x, h1 = pb.w - pb.y + eps;
            const float w2 = tb.z - tb.x, h2 = tb.w - tb.y + eps;
            const float iw_raw = fminf(pb.z, tb.z) - fmaxf(pb.x, tb.x);
            const float ih_raw = fminf(pb.w, tb.w) - fmaxf(pb.y, tb.y);
            const float iw = fmaxf(iw_raw, 0.f), ih = fmaxf(ih_raw, 0.f);
            const float inter = iw * ih;
            const float uni = w1 * h1 + w2 * h2 - inter + eps;
            const float iou = inter / uni;
            const float cw = fmaxf(pb.z, tb.z) - fminf(pb.x, tb.x);
            const float ch = fmaxf(pb.w, tb.w) - fminf(pb.y, tb.y);
            const float c2 = cw * cw + ch * ch + eps;
            const float dx = tb.x + tb.z - pb.x - pb.z, dy = tb.y + tb.w - pb.y - pb.w;
            const float rho2 = (dx * dx + dy * dy) * 0.25f;
            const float kpi = 4.f / (3.14159265358979323846f * 3.14159265358979323846f);
            const float da = atanf(w2 / h2) - atanf(w1 / h1);
            const float v = kpi * da * da;
            const float alpha = v / (v - iou + (1.f + eps));
            const float ciou = iou - (rho2 / c2 + v * alpha);
            l_iou = (1.f - ciou) * w * inv_tss;
            // ---- gradient of ciou w.r.t. pred box (x1,y1,x2,y2) ----
            // intersection extents
            const float diw[4] = {(iw_raw > 0.f && pb.x > tb.x) ? -1.f : 0.f, 0.f, (iw_raw > 0.f && pb.z < tb.z) ? 1.f : 0.f, 0.f};
            const float dih[4] = {0.f, (ih_raw > 0.f && pb.y > tb.y) ? -1.f : 0.f, 0.f, (ih_raw > 0.f && pb.w < tb.w) ? 1.f : 0.f};
            const float dw1[4] = {-1.f, 0.f, 1.f, 0.f}, dh1[4] = {0.f, -1.f, 0.f, 1.f};
            const float dcw[4] = {(pb.x < tb.x) ? -1.f : 0.f, 0.f, (pb.z > tb.z) ? 1.f : 0.f, 0.f};
            const float dch[4] = {0.f, (pb.y < tb.y) ? -1.f : 0.f, 0.f, (pb.w > tb.w) ? 1.f : 0.f};
            const float drx[4] = {-1.f, 0.f, -1.f, 0.f}, dry[4] = {0.f, -1.f, 0.f, -1.f};
            const float q = w1 * w1 + h1 * h1;
            const float coef = -w * inv_tss;  // d loss / d ciou
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float dinter = diw[k] * ih + iw * dih[k];
                const float duni = dw1[k] * h1 + w1 * dh1[k] - dinter;
                const float diou = (dinter * uni - inter * duni) / (uni * uni);
                const float dc2 = 2.f * cw * dcw[k] + 2.f * ch * dch[k];
                const float drho2 = 0.5f * (dx * drx[k] + dy * dry[k]);
                const float dpen = (drho2 * c2 - rho2 * dc2) / (c2 * c2);
                // d atan(w1/h1) = (h1 dw1 - w1 dh1)/(w1^2+h1^2);  v = k (A2-A1)^2
                const float dA1 = (h1 * dw1[k] - w1 * dh1[k]) / q;
                const float dv = -2.f * kpi * da * dA1;
                gb[k] = coef * (diou - dpen - alpha * dv);
            }
            // ---- DFL (utils/loss.py:77-87) ----
            if (p.use_dfl) {
                const float ax = p.anchor_points[(i % p.A) * 2], ay = p.anchor_points[(i % p.A) * 2 + 1];
                const float hi = (float)p.reg_max - 0.01f;
                const float tgt[4] = {fminf(fmaxf(ax - tb.x, 0.f), hi), fminf(fmaxf(ay - tb.y, 0.f), hi),
                                      fminf(fmaxf(tb.z - ax, 0.f), hi), fminf(fmaxf(tb.w - ay, 0.f), hi)};
                const float* pd = p.pred_dist + i * 4 * bins;
                float side_sum = 0.f;
                for (int s = 0; s < 4; ++s) {
                    const int tl = (int)tgt[s], tr = tl + 1;
                    const float wl = (float)tr - tgt[s], wr = 1.f - wl;
                    const float* lg = pd + s * bins;
                    float mx = -3.4e38f;
                    for (int j = 0; j < bins; ++j) mx = fmaxf(mx, lg[j]);
                    float se = 0.f;
                    for (int j = 0; j < bins; ++j) se += expf(lg[j] - mx);
                    const float lse = mx + logf(se);
                    side_sum += (lse - lg[tl]) * wl + (lse - lg[tr]) * wr;
                    if (gd) {
                        const float sc = 0.25f * w * inv_tss;
                        for (int j = 0; j < bins; ++j) {
                            float g = expf(lg[j] - lse);
                            if (j == tl) g -= wl;
                            if (j == tr) g -= wr;
                            gd[s * bins + j] = g * sc;
                        }
                    }
                }
                l_dfl = side_sum * 0.25f * w * inv_tss;
            } else if (gd) {
                for (int j = 0; j < 4 * bins; ++j) gd[j] = 0.f;
            }
        } else if (gd) {
            for (int j = 0; j < 4 * bins; ++j) gd[j] = 0.f;
        }
        if (p.grad_bboxes) reinterpret_cast<float4*>(p.grad_bboxes)[i] = make_float4(gb[0], gb[1], gb[2], gb[3]);
    }
    l_iou = warp_sum(l_iou);
    l_dfl = warp_sum(l_dfl);
    if ((threadIdx.x & 31) == 0) { s_red[0][threadIdx.x >> 5] = l_iou; s_red[1][threadIdx.x >> 5] = l_dfl; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, b = 0.f;
        for (int k = 0; k < BL_NT / 32; ++k) { a += s_red[0][k]; b += s_red[1][k]; }
        p.part[blockIdx.x] = a;
        p.part[gridDim.x + blockIdx.x] = b;
    }
}

__global__ void __launch_bounds__(1024) k_bbox_loss_finalize(const float* part, int nblk, float* out) {
    __shared__ double s[2][32];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    double a = 0.0, b = 0.0;
    for (int i = tid; i < nblk; i += 1024) { a += (double)part[i]; b += (double)part[nblk + i]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
    if (lane == 0) { s[0][warp] = a; s[1][warp] = b; }
    __syncthreads();
    if (tid == 0) {
        double ta = 0.0, tb = 0.0;
        for (int w = 0; w < 32; ++w) { ta += s[0][w]; tb += s[1][w]; }
        out[0] = (float)ta;
        out[1] = (float)tb;
    }
}

size_t bbox_loss_workspace_bytes(int B, int A) {
    const int64_t n = (int64_t)B * A;
    return align_up((size_t)(2 * ((n + BL_NT - 1) / BL_NT) + 2) * sizeof(float), 256);
}

int launch_bbox_loss(const float* pred_dist, const float* pred_bboxes, const float* anchor_points,
                     const float* target_bboxes, const float* target_scores, const uint8_t* fg_mask, const float* tss_d,
                     int B, int A, int nc, int reg_max, int use_dfl, float* loss_out, float* grad_dist, float* grad_bboxes,
                     void* workspace, size_t workspace_bytes, cudaStream_t st) {
    if (bbox_loss_workspace_bytes(B, A) > workspace_bytes) { ycr_set_error("bbox loss workspace too small"); return YCR_E_WORKSPACE; }
    const int64_t n = (int64_t)B * A;
    const int nblk = (int)((n + BL_NT - 1) / BL_NT);
    BboxLossArgs p{pred_dist, pred_bboxes, anchor_points, target_bboxes, target_scores, fg_mask, tss_d, grad_dist, grad_bboxes,
                   reinterpret_cast<float*>(workspace), B, A, nc, reg_max, use_dfl};
    k_bbox_loss<<<nblk, BL_NT, 0, st>>>(p);
    YCR_LAUNCH_CHECK();
    k_bbox_loss_finalize<<<1, 1024, 0, st>>>(reinterpret_cast<float*>(workspace), nblk, loss_out);
    YCR_LAUNCH_CHECK();
    return YCR_OK;
}
