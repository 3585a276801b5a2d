// extern "C" surface declared in include/ycr_b200.h.  Thin argument checking + kernel sequencing;
// no torch types, no CPU fallbacks.
#include "train_path.cuh"
#include "dtype.cuh"
#include <stdarg.h>
#include <string.h>

static thread_local char g_err[512] = "";

void ycr_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ---- per-kernel event timing --------------------------------------------------------------------
#include <thread>
#include <vector>
#include <algorithm>
static bool g_prof_on = false;
static std::vector<cudaEvent_t> g_prof_ev;       // pool
static std::vector<int> g_prof_tag;              // tag of pair k (events 2k, 2k+1)
static size_t g_prof_used = 0;
static unsigned g_prof_mask = 0xFFFFFFFFu;

void ycr_prof_mark(int tag, int end, cudaStream_t st) {
    if (!g_prof_on || !((g_prof_mask >> tag) & 1u)) return;
    if (!end) {
        if (g_prof_used + 2 > g_prof_ev.size()) return;
        g_prof_tag.push_back(tag);
        cudaEventRecord(g_prof_ev[g_prof_used], st);
    } else {
        if (g_prof_tag.size() * 2 != g_prof_used + 2) return;
        cudaEventRecord(g_prof_ev[g_prof_used + 1], st);
        g_prof_used += 2;
    }
}

int debug_stats(unsigned long long* out_h, int reset);
int launch_resample(const float* pts, const int* offsets, int S, int n_out, float* out, cudaStream_t st);
size_t bbox_loss_workspace_bytes(int B, int A);
int launch_bbox_loss(const float* pred_dist, const float* pred_bboxes, const float* anchor_points,
                     const float* target_bboxes, const float* target_scores, const uint8_t* fg_mask, const float* tss_d,
                     int B, int A, int nc, int reg_max, int use_dfl, float* loss_out, float* grad_dist, float* grad_bboxes,
                     void* workspace, size_t workspace_bytes, cudaStream_t st);
int launch_scale(void* const* p, const int64_t* n, int n_levels, int dtype, const float* scale, cudaStream_t st);
int launch_pack_targets(const float* head, int64_t hs, const float* seg, int64_t ss, int N, int B, int G, float img_w, float img_h,
                        float* out, cudaStream_t st);
int launch_decode(const ycr_grid_t* grid, const void* const* feats, int dtype, int B, int nc, int R, float* allpred, int2* best,
                  cudaStream_t st);
size_t nms_workspace_bytes(int B, int A, const ycr_nms_cfg_t* cfg);
int launch_nms(const float* prediction, int B, int CH, int A, const ycr_nms_cfg_t* cfg, float* out_rows, int* out_counts,
               void* workspace, size_t workspace_bytes, cudaStream_t st);

int launch_rasterize(const float* rows, int64_t row_stride, int n, int R, int H, int W, uint8_t* masks, cudaStream_t st);
size_t mask_iou_workspace_bytes(int N, int M, int64_t n);
int launch_mask_iou(const void* m1, int dt1, const void* m2, int dt2, int N, int M, int64_t n, float eps, float* iou,
                    void* workspace, size_t workspace_bytes, cudaStream_t st);

int launch_pack_targets_mapped(const float* head, int64_t hs, const float* seg, int64_t ss, const int* row_of, int B, int G,
                               float img_w, float img_h, float* out, cudaStream_t st);

size_t detect_workspace_bytes(int B, int A, const ycr_nms_cfg_t* cfg);
int launch_detect(const ycr_grid_t* grid, const void* const* feats, int dtype, int B, int nc, int R, const ycr_nms_cfg_t* cfg,
                  float* out_rows, int* out_counts, void* workspace, size_t workspace_bytes, cudaStream_t st);

static int check_common(const ycr_grid_t* grid, const ycr_assign_cfg_t* cfg, int B, int G) {
    if (!grid || !cfg) { ycr_set_error("null grid/cfg"); return YCR_E_ARG; }
    if (grid->n_levels < 1 || grid->n_levels > YCR_MAX_LEVELS) { ycr_set_error("n_levels %d out of range", grid->n_levels); return YCR_E_ARG; }
    if (cfg->rays != 36 && cfg->rays != 72) { ycr_set_error("rays must be 36 or 72, got %d", cfg->rays); return YCR_E_ARG; }
    if (cfg->topk < 1 || cfg->topk > 64) { ycr_set_error("topk %d out of range [1,64]", cfg->topk); return YCR_E_ARG; }
    if (B < 1 || G < 0 || G > 65535) { ycr_set_error("B=%d G=%d out of range", B, G); return YCR_E_ARG; }
    {
        // the per-image resolution kernel keeps the image's anchors, GT descriptors and positives in shared memory
        int64_t A = 0;
        for (int l = 0; l < grid->n_levels; ++l) A += (int64_t)grid->h[l] * grid->w[l];
        const int64_t need = ycr_resolve_smem_bytes(A, G, cfg->topk);
        if (need > YCR_SMEM_MAX) {
            ycr_set_error("%d GTs per image with %lld anchors and topk %d need %lld bytes of shared memory per image "
                          "(limit %d): lower the number of instances per image", G, (long long)A, cfg->topk, (long long)need,
                          YCR_SMEM_MAX);
            return YCR_E_ARG;
        }
    }
    return YCR_OK;
}

extern "C" {

const char* ycr_last_error(void) { return g_err; }

int ycr_profile_begin(int max_records) {
    if (max_records < 1) { ycr_set_error("max_records must be positive"); return YCR_E_ARG; }
    while ((int)g_prof_ev.size() < 2 * max_records) {
        cudaEvent_t e;
        YCR_CUDA_CHECK(cudaEventCreate(&e));
        g_prof_ev.push_back(e);
    }
    g_prof_tag.clear();
    g_prof_used = 0;
    g_prof_on = true;
    return YCR_OK;
}

int ycr_profile_select(unsigned tag_mask) {
    g_prof_mask = tag_mask;
    return YCR_OK;
}

int ycr_debug_stats(unsigned long long* out_h, int reset) {
    if (!out_h) { ycr_set_error("null argument"); return YCR_E_ARG; }
    return debug_stats(out_h, reset);
}

int ycr_profile_end(float* ms_sum, int* count) {
    g_prof_on = false;
    if (!ms_sum || !count) { ycr_set_error("null argument"); return YCR_E_ARG; }
    for (int t = 0; t < YCR_T_COUNT; ++t) { ms_sum[t] = 0.f; count[t] = 0; }
    const size_t pairs = g_prof_used / 2;
    for (size_t k = 0; k < pairs; ++k) {
        YCR_CUDA_CHECK(cudaEventSynchronize(g_prof_ev[2 * k + 1]));
        float ms = 0.f;
        YCR_CUDA_CHECK(cudaEventElapsedTime(&ms, g_prof_ev[2 * k], g_prof_ev[2 * k + 1]));
        const int t = g_prof_tag[k];
        if (t >= 0 && t < YCR_T_COUNT) { ms_sum[t] += ms; count[t] += 1; }
    }
    g_prof_tag.clear();
    g_prof_used = 0;
    return YCR_OK;
}
int ycr_version(void) { return 100; }

int ycr_abi_sizes(int* sizes_out) {
    if (!sizes_out) { ycr_set_error("null argument"); return YCR_E_ARG; }
    const size_t s[7] = {sizeof(ycr_grid_t), sizeof(ycr_pred_view_t), sizeof(ycr_gt_t), sizeof(ycr_assign_cfg_t),
                         sizeof(ycr_assign_out_t), sizeof(ycr_loss_cfg_t), sizeof(ycr_nms_cfg_t)};
    for (int i = 0; i < 7; ++i) sizes_out[i] = (int)s[i];
    return 7;
}

int64_t ycr_candidate_bound_h(const ycr_grid_t* grid, const float* boxes_h, int64_t row_stride, int n_rows) {
    int64_t total = 0;
    for (int r = 0; r < n_rows; ++r) {
        const float* bx = boxes_h + r * row_stride;
        const double w = (double)bx[2] - bx[0], h = (double)bx[3] - bx[1];
        if (!(w > 0) || !(h > 0)) continue;
        for (int l = 0; l < grid->n_levels; ++l) {
            int64_t cx = (int64_t)floor(w / grid->stride[l]) + 2, cy = (int64_t)floor(h / grid->stride[l]) + 2;
            if (cx > grid->w[l]) cx = grid->w[l];
            if (cy > grid->h[l]) cy = grid->h[l];
            total += cx * cy;
        }
    }
    return total;
}

int64_t ycr_candidate_bound_xywhn_h(const ycr_grid_t* grid, const float* xywhn_h, int64_t row_stride, int n_rows, float img_w,
                                    float img_h) {
    int64_t total = 0;
    for (int r = 0; r < n_rows; ++r) {
        const float* bx = xywhn_h + r * row_stride;
        const double w = (double)bx[2] * img_w, h = (double)bx[3] * img_h;
        if (!(w > 0) || !(h > 0)) continue;
        for (int l = 0; l < grid->n_levels; ++l) {
            int64_t cx = (int64_t)floor(w / grid->stride[l]) + 2, cy = (int64_t)floor(h / grid->stride[l]) + 2;
            if (cx > grid->w[l]) cx = grid->w[l];
            if (cy > grid->h[l]) cy = grid->h[l];
            total += cx * cy;
        }
    }
    return total;
}

int ycr_stage_targets_h(const float* batch_idx_h, const float* cls_h, const float* bboxes_h, const float* const* seg_ptrs_h,
                        const int* seg_rows_h, int n_seg_tensors, int N, int B, const ycr_grid_t* grid, float img_w, float img_h,
                        float* staging_h, int* G_out, int64_t* cand_bound_out) {
    if (!batch_idx_h || !cls_h || !bboxes_h || !seg_ptrs_h || !seg_rows_h || !grid || !staging_h || !G_out || !cand_bound_out ||
        N < 0 || B < 1) {
        ycr_set_error("bad stage_targets arguments");
        return YCR_E_ARG;
    }
    float* head = staging_h;
    float* seg = staging_h + (size_t)N * 6;
    int64_t rows = 0;
    std::vector<int64_t> first((size_t)n_seg_tensors + 1, 0);
    for (int k = 0; k < n_seg_tensors; ++k) {
        if (seg_rows_h[k] < 0 || rows + seg_rows_h[k] > N) { ycr_set_error("segment rows do not add up to the %d boxes", N); return YCR_E_ARG; }
        first[(size_t)k] = rows;
        rows += seg_rows_h[k];
    }
    first[(size_t)n_seg_tensors] = rows;
    if (rows != N) { ycr_set_error("segment rows (%lld) do not match the %d boxes", (long long)rows, N); return YCR_E_ARG; }
    // the one pass over the contours (C2: 3.7 MB, a quarter of a millisecond on one core - most of this call).
    // YCR_STAGE_THREADS=n splits it over n-1 short-lived helper threads (4: 230 -> 115 us on the GPU box's host); off by
    // default: the path is device-bound at batch 64, and with the helpers the step measured 2.5 % SLOWER (1.240 vs
    // 1.210 ms, twice each on one box)
    auto copy_range = [&](int k0, int k1) {
        for (int k = k0; k < k1; ++k)
            memcpy(seg + (size_t)first[(size_t)k] * 2 * YCR_C, seg_ptrs_h[k], (size_t)seg_rows_h[k] * 2 * YCR_C * sizeof(float));
    };
    const size_t bytes = (size_t)N * 2 * YCR_C * sizeof(float);
    int nthr = 1;
    if (bytes >= ((size_t)1 << 20) && n_seg_tensors >= 8) {
        static const int hw = [] {
            const char* e = getenv("YCR_STAGE_THREADS");
            const int want = e ? atoi(e) : 1;
            const int have = (int)std::thread::hardware_concurrency();
            return want < 1 ? 1 : (have > 0 && want > have ? have : want);
        }();
        nthr = hw;
    }
    if (nthr <= 1) {
        copy_range(0, n_seg_tensors);
    } else {
        std::vector<std::thread> helpers;
        helpers.reserve((size_t)nthr - 1);
        for (int t = 1; t < nthr; ++t)
            helpers.emplace_back(copy_range, (int)((int64_t)n_seg_tensors * t / nthr), (int)((int64_t)n_seg_tensors * (t + 1) / nthr));
        copy_range(0, n_seg_tensors / nthr);
        for (auto& h : helpers) h.join();
    }
    std::vector<int> count((size_t)B, 0);
    int G = 0;
    for (int n = 0; n < N; ++n) {
        float* h = head + (size_t)n * 6;
        h[0] = batch_idx_h[n];
        h[1] = cls_h[n];
        h[2] = bboxes_h[4 * n]; h[3] = bboxes_h[4 * n + 1]; h[4] = bboxes_h[4 * n + 2]; h[5] = bboxes_h[4 * n + 3];
        const int b = (int)batch_idx_h[n];
        if (b >= 0 && b < B) { const int c = ++count[(size_t)b]; if (c > G) G = c; }
    }
    *G_out = G;
    *cand_bound_out = ycr_candidate_bound_xywhn_h(grid, head + 2, 6, N, img_w, img_h);
    // row of every (image, slot): slot = rank of the row among its image's rows, in row order (utils/loss.py:228-232)
    int* row_of = reinterpret_cast<int*>(staging_h + (size_t)N * (6 + 2 * YCR_C));
    for (size_t i = 0; i < (size_t)B * G; ++i) row_of[i] = -1;
    std::fill(count.begin(), count.end(), 0);
    for (int n = 0; n < N; ++n) {
        const int b = (int)batch_idx_h[n];
        if (b >= 0 && b < B) row_of[(size_t)b * G + count[(size_t)b]++] = n;
    }
    return YCR_OK;
}

int ycr_pack_targets_mapped(const float* head, int64_t head_stride, const float* segments, int64_t seg_stride, const int* row_of,
                            int B, int G, float img_w, float img_h, float* out_packed, void* stream) {
    if (!out_packed || (B * G > 0 && (!head || !segments || !row_of))) { ycr_set_error("null argument"); return YCR_E_ARG; }
    if (B < 1 || G < 0 || head_stride < 6 || seg_stride < 2 * YCR_C) { ycr_set_error("bad B/G/strides"); return YCR_E_ARG; }
    return launch_pack_targets_mapped(head, head_stride, segments, seg_stride, row_of, B, G, img_w, img_h, out_packed,
                                      reinterpret_cast<cudaStream_t>(stream));
}

size_t ycr_assign_workspace_bytes(const ycr_grid_t* grid, int B, int G, const ycr_assign_cfg_t* cfg, int64_t cand_capacity) {
    if (check_common(grid, cfg, B, G)) return 0;
    GridDev gd = make_grid_dev(grid);
    return assign_ws_layout(nullptr, nullptr, gd, B, G, cfg->topk, cfg->rays, cand_capacity, false);
}

size_t ycr_seg_loss_workspace_bytes(const ycr_grid_t* grid, int B, int G, const ycr_assign_cfg_t* cfg, int64_t cand_capacity) {
    if (check_common(grid, cfg, B, G)) return 0;
    GridDev gd = make_grid_dev(grid);
    return assign_ws_layout(nullptr, nullptr, gd, B, G, cfg->topk, cfg->rays, cand_capacity, true);
}

int ycr_assign(const ycr_grid_t* grid, const ycr_pred_view_t* pred, const ycr_gt_t* gt, const ycr_assign_cfg_t* cfg,
               const ycr_assign_out_t* out, void* workspace, size_t workspace_bytes, int64_t cand_capacity, void* stream) {
    if (!pred || !gt || !out || !workspace) { ycr_set_error("null argument"); return YCR_E_ARG; }
    int rc = check_common(grid, cfg, gt->B, gt->G);
    if (rc) return rc;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    AssignArgs a{};
    a.grid = make_grid_dev(grid);
    a.pred = *pred;
    a.gt = *gt;
    a.cfg = *cfg;
    a.pc = make_polar_const(cfg->rays);
    AssignWs ws;
    const size_t need = assign_ws_layout(&ws, workspace, a.grid, gt->B, gt->G, cfg->topk, cfg->rays, cand_capacity, false);
    if (need > workspace_bytes) { ycr_set_error("assign workspace too small: need %zu have %zu", need, workspace_bytes); return YCR_E_WORKSPACE; }
    if ((rc = launch_assign_core(a, ws, out->n_pos_d, st))) return rc;
    if ((rc = launch_positive_targets(a, ws, out->gt_dist, out->centerness, out->pos_capacity, false, nullptr, st))) return rc;
    return launch_assign_dense(a, ws, *out, st);
}

int ycr_seg_loss_fwd_bwd_dt(const ycr_grid_t* grid, const void* const* feats, void* const* grad_feats, int dtype,
                            const ycr_gt_t* gt, const ycr_assign_cfg_t* acfg, const ycr_loss_cfg_t* lcfg, float* loss_out,
                            void* workspace, size_t workspace_bytes, int64_t cand_capacity, void* stream) {
    if (!feats || !gt || !lcfg || !loss_out || !workspace) { ycr_set_error("null argument"); return YCR_E_ARG; }
    if (dtype != YCR_F32 && dtype != YCR_F16 && dtype != YCR_BF16) { ycr_set_error("dtype %d: 0 f32, 1 f16, 2 bf16", dtype); return YCR_E_ARG; }
    int rc = check_common(grid, acfg, gt->B, gt->G);
    if (rc) return rc;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    AssignArgs a{};
    a.grid = make_grid_dev(grid);
    a.gt = *gt;
    a.cfg = *acfg;
    a.pc = make_polar_const(acfg->rays);
    a.dtype = dtype;
    const int R = acfg->rays, nc = acfg->num_classes;
    const int esz = ycr_dtype_size(dtype);
    for (int l = 0; l < grid->n_levels; ++l) {
        const int64_t hw = (int64_t)grid->h[l] * grid->w[l];
        // (the view's pointers are typed float*; the kernels index them as arrays of `dtype` elements)
        a.pred.rays[l] = reinterpret_cast<const float*>(feats[l]);
        a.pred.cls[l] = reinterpret_cast<const float*>(reinterpret_cast<const char*>(feats[l]) + (int64_t)R * hw * esz);
        a.pred.rays_sb[l] = a.pred.cls_sb[l] = (int64_t)(R + nc) * hw;
        a.pred.rays_sa[l] = a.pred.cls_sa[l] = 1;
        a.pred.rays_sc[l] = a.pred.cls_sc[l] = hw;
        a.pred.ray_scale[l] = grid->stride[l];
    }
    a.pred.cls_is_logit = 1;
    AssignWs ws;
    const size_t need = assign_ws_layout(&ws, workspace, a.grid, gt->B, gt->G, acfg->topk, R, cand_capacity, true);
    if (need > workspace_bytes) { ycr_set_error("loss workspace too small: need %zu have %zu", need, workspace_bytes); return YCR_E_WORKSPACE; }
    if ((rc = launch_assign_core(a, ws, nullptr, st))) return rc;
    if ((rc = launch_positive_targets(a, ws, nullptr, nullptr, 0, true, lcfg, st))) return rc;
    return launch_loss_stream(a, ws, feats, grad_feats, *lcfg, loss_out, st);
}

int ycr_seg_loss_fwd_bwd(const ycr_grid_t* grid, const float* const* feats, float* const* grad_feats, const ycr_gt_t* gt,
                         const ycr_assign_cfg_t* acfg, const ycr_loss_cfg_t* lcfg, float* loss_out, void* workspace,
                         size_t workspace_bytes, int64_t cand_capacity, void* stream) {
    return ycr_seg_loss_fwd_bwd_dt(grid, reinterpret_cast<const void* const*>(feats), reinterpret_cast<void* const*>(grad_feats),
                                   YCR_F32, gt, acfg, lcfg, loss_out, workspace, workspace_bytes, cand_capacity, stream);
}

int ycr_scale_grads_dt(const ycr_grid_t* grid, int B, int channels, void* const* grad_feats, int dtype, const float* scale_d,
                       void* stream) {
    if (!grid || !grad_feats || !scale_d) { ycr_set_error("null argument"); return YCR_E_ARG; }
    if (dtype != YCR_F32 && dtype != YCR_F16 && dtype != YCR_BF16) { ycr_set_error("dtype %d: 0 f32, 1 f16, 2 bf16", dtype); return YCR_E_ARG; }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    int64_t n[YCR_MAX_LEVELS] = {0};
    if (grid->n_levels > YCR_MAX_LEVELS) { ycr_set_error("too many levels"); return YCR_E_ARG; }
    for (int l = 0; l < grid->n_levels; ++l) n[l] = (int64_t)B * channels * grid->h[l] * grid->w[l];
    return launch_scale(grad_feats, n, grid->n_levels, dtype, scale_d, st);
}

int ycr_scale_grads(const ycr_grid_t* grid, int B, int channels, float* const* grad_feats, const float* scale_d, void* stream) {
    return ycr_scale_grads_dt(grid, B, channels, reinterpret_cast<void* const*>(grad_feats), YCR_F32, scale_d, stream);
}

int ycr_pack_targets(const float* targets, int64_t row_stride, int N, int B, int G, float img_w, float img_h,
                     float* out_packed, void* stream) {
    if (!out_packed || (N > 0 && !targets)) { ycr_set_error("null argument"); return YCR_E_ARG; }
    if (B < 1 || G < 0 || row_stride < 6 + 2 * YCR_C) { ycr_set_error("bad B/G/row_stride"); return YCR_E_ARG; }
    if (G == 0) return YCR_OK;
    return launch_pack_targets(targets, row_stride, targets + 6, row_stride, N, B, G, img_w, img_h, out_packed,
                               reinterpret_cast<cudaStream_t>(stream));
}

int ycr_pack_targets_split(const float* head, int64_t head_stride, const float* segments, int64_t seg_stride, int N, int B, int G,
                           float img_w, float img_h, float* out_packed, void* stream) {
    if (!out_packed || (N > 0 && (!head || !segments))) { ycr_set_error("null argument"); return YCR_E_ARG; }
    if (B < 1 || G < 0 || head_stride < 6 || seg_stride < 2 * YCR_C) { ycr_set_error("bad B/G/strides"); return YCR_E_ARG; }
    if (G == 0) return YCR_OK;
    return launch_pack_targets(head, head_stride, segments, seg_stride, N, B, G, img_w, img_h, out_packed,
                               reinterpret_cast<cudaStream_t>(stream));
}

int ycr_resample_segments(const float* pts, const int* offsets, int S, int n_out, float* out, void* stream) {
    if (S < 0 || n_out < 2 || (S > 0 && (!pts || !offsets || !out))) { ycr_set_error("bad resample arguments"); return YCR_E_ARG; }
    return launch_resample(pts, offsets, S, n_out, out, reinterpret_cast<cudaStream_t>(stream));
}

size_t ycr_bbox_loss_workspace_bytes(int B, int A) { return (B > 0 && A > 0) ? bbox_loss_workspace_bytes(B, A) : 0; }

int ycr_bbox_loss_fwd_bwd(const float* pred_dist, const float* pred_bboxes, const float* anchor_points,
                          const float* target_bboxes, const float* target_scores, const uint8_t* fg_mask,
                          const float* target_scores_sum_d, int B, int A, int nc, int reg_max, int use_dfl,
                          float* loss_out, float* grad_pred_dist, float* grad_pred_bboxes, void* workspace,
                          size_t workspace_bytes, void* stream) {
    if (!pred_bboxes || !target_bboxes || !target_scores || !fg_mask || !target_scores_sum_d || !loss_out || !workspace ||
        (use_dfl && (!pred_dist || !anchor_points))) { ycr_set_error("null argument"); return YCR_E_ARG; }
    if (B < 1 || A < 1 || nc < 1 || reg_max < 1 || reg_max > 63) { ycr_set_error("bad B/A/nc/reg_max"); return YCR_E_ARG; }
    return launch_bbox_loss(pred_dist, pred_bboxes, anchor_points, target_bboxes, target_scores, fg_mask, target_scores_sum_d,
                            B, A, nc, reg_max, use_dfl, loss_out, grad_pred_dist, grad_pred_bboxes, workspace,
                            workspace_bytes, reinterpret_cast<cudaStream_t>(stream));
}

int ycr_decode(const ycr_grid_t* grid, const float* const* feats, int B, int nc, int R, float* allpred, void* stream) {
    if (!grid || !feats || !allpred) { ycr_set_error("null argument"); return YCR_E_ARG; }
    if (R < 1 || R > 72 || 360 % R) { ycr_set_error("unsupported R=%d", R); return YCR_E_ARG; }
    return launch_decode(grid, reinterpret_cast<const void* const*>(feats), YCR_F32, B, nc, R, allpred, nullptr,
                         reinterpret_cast<cudaStream_t>(stream));
}

int ycr_decode_best(const ycr_grid_t* grid, const float* const* feats, int B, int nc, int R, float* allpred,
                    void* best_class_out, void* stream) {
    if (!grid || !feats || !allpred || !best_class_out) { ycr_set_error("null argument"); return YCR_E_ARG; }
    if (R < 1 || R > 72 || 360 % R) { ycr_set_error("unsupported R=%d", R); return YCR_E_ARG; }
    return launch_decode(grid, reinterpret_cast<const void* const*>(feats), YCR_F32, B, nc, R, allpred,
                         reinterpret_cast<int2*>(best_class_out), reinterpret_cast<cudaStream_t>(stream));
}

int ycr_decode_dt(const ycr_grid_t* grid, const void* const* feats, int dtype, int B, int nc, int R, float* allpred,
                  void* best_class_out, void* stream) {
    if (!grid || !feats || !allpred) { ycr_set_error("null argument"); return YCR_E_ARG; }
    if (R < 1 || R > 72 || 360 % R) { ycr_set_error("unsupported R=%d", R); return YCR_E_ARG; }
    if (dtype != YCR_F32 && dtype != YCR_F16 && dtype != YCR_BF16) { ycr_set_error("dtype %d: 0 f32, 1 f16, 2 bf16", dtype); return YCR_E_ARG; }
    return launch_decode(grid, feats, dtype, B, nc, R, allpred, reinterpret_cast<int2*>(best_class_out),
                         reinterpret_cast<cudaStream_t>(stream));
}

size_t ycr_nms_workspace_bytes(int B, int A, int channels, const ycr_nms_cfg_t* cfg) {
    (void)channels;
    if (!cfg) return 0;
    return nms_workspace_bytes(B, A, cfg);
}

int ycr_nms(const float* prediction, int B, int channels, int A, const ycr_nms_cfg_t* cfg, float* out_rows, int* out_counts,
            void* workspace, size_t workspace_bytes, void* stream) {
    if (!prediction || !cfg || !out_rows || !out_counts || !workspace) { ycr_set_error("null argument"); return YCR_E_ARG; }
    if (cfg->conf_thres < 0.f || cfg->conf_thres > 1.f || cfg->iou_thres < 0.f || cfg->iou_thres > 1.f) {
        ycr_set_error("thresholds must lie in [0,1]");
        return YCR_E_ARG;
    }
    if (cfg->nc < 1 || 4 + cfg->nc > channels) { ycr_set_error("bad nc %d for %d channels", cfg->nc, channels); return YCR_E_ARG; }
    return launch_nms(prediction, B, channels, A, cfg, out_rows, out_counts, workspace, workspace_bytes,
                      reinterpret_cast<cudaStream_t>(stream));
}

int ycr_rasterize_contours(const float* rows, int64_t row_stride, int n, int R, int H, int W, uint8_t* masks, void* stream) {
    if (n < 0 || R < 1 || R > 72 || H < 1 || W < 1 || row_stride < 6 + 3 * R || (n > 0 && (!rows || !masks))) {
        ycr_set_error("bad rasterize arguments");
        return YCR_E_ARG;
    }
    return launch_rasterize(rows, row_stride, n, R, H, W, masks, reinterpret_cast<cudaStream_t>(stream));
}

size_t ycr_mask_iou_workspace_bytes(int N, int M, int64_t n) { return (N < 0 || M < 0 || n < 1) ? 0 : mask_iou_workspace_bytes(N, M, n); }

int ycr_mask_iou(const void* mask1, int dtype1, const void* mask2, int dtype2, int N, int M, int64_t n, float eps, float* iou,
                 void* workspace, size_t workspace_bytes, void* stream) {
    if (N < 0 || M < 0 || n < 1 || ((N > 0 && M > 0) && (!mask1 || !mask2 || !iou || !workspace))) {
        ycr_set_error("bad mask_iou arguments");
        return YCR_E_ARG;
    }
    return launch_mask_iou(mask1, dtype1, mask2, dtype2, N, M, n, eps, iou, workspace, workspace_bytes,
                           reinterpret_cast<cudaStream_t>(stream));
}

size_t ycr_detect_workspace_bytes(const ycr_grid_t* grid, int B, const ycr_nms_cfg_t* cfg) {
    if (!grid || !cfg || B < 1 || grid->n_levels < 1 || grid->n_levels > YCR_MAX_LEVELS) return 0;
    int A = 0;
    for (int l = 0; l < grid->n_levels; ++l) A += grid->h[l] * grid->w[l];
    return detect_workspace_bytes(B, A, cfg);
}

int ycr_detect(const ycr_grid_t* grid, const void* const* feats, int dtype, int B, int nc, int R, const ycr_nms_cfg_t* cfg,
               float* out_rows, int* out_counts, void* workspace, size_t workspace_bytes, void* stream) {
    if (!grid || !feats || !cfg || !out_rows || !out_counts || !workspace) { ycr_set_error("null argument"); return YCR_E_ARG; }
    if (grid->n_levels < 1 || grid->n_levels > YCR_MAX_LEVELS || B < 1 || nc < 1) { ycr_set_error("bad grid / B / nc"); return YCR_E_ARG; }
    if (R < 1 || R > 72 || 360 % R) { ycr_set_error("unsupported R=%d", R); return YCR_E_ARG; }
    if (dtype != YCR_F32 && dtype != YCR_F16 && dtype != YCR_BF16) { ycr_set_error("dtype %d: 0 f32, 1 f16, 2 bf16", dtype); return YCR_E_ARG; }
    if (cfg->conf_thres < 0.f || cfg->conf_thres > 1.f || cfg->iou_thres < 0.f || cfg->iou_thres > 1.f) {
        ycr_set_error("thresholds must lie in [0,1]");
        return YCR_E_ARG;
    }
    return launch_detect(grid, feats, dtype, B, nc, R, cfg, out_rows, out_counts, workspace, workspace_bytes,
                         reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
