/* ycr_b200.h — C ABI of the B200-native polar-contour hot path.
 *
 * The reference (ai4in/YOLO-Contour-Regression) is pure Python/PyTorch and has no FFI layer; its
 * boundary for this path is the Python symbol surface listed in SURVEY.md §8(b).  Each entry point
 * below replaces one of those symbols (cited as file:line under
 * /root/reference/ultralytics-main/ultralytics/) and is what a ctypes / cffi / TORCH_LIBRARY stub on
 * the reference side binds (INTEGRATION.md shows the stubs).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless the name ends in _h;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); nothing synchronises the
 *     host unless stated; inputs are never written;
 *   - fp32 data, int32 indices unless stated; the *_i64 outputs exist because the reference API
 *     returns int64 tensors;
 *   - return value: 0 on success, a negative YCR_E_* code otherwise (never a CPU fallback);
 *   - workspace is caller-provided; size it with the matching *_workspace_bytes() call.
 */
#ifndef YCR_B200_H
#define YCR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define YCR_MAX_LEVELS 4
#define YCR_CONTOUR_POINTS 360 /* fixed by the reference wire format, utils/instance.py:202 */

#define YCR_OK 0
#define YCR_E_ARG (-1)       /* bad argument (shape, null pointer, unsupported R) */
#define YCR_E_WORKSPACE (-2) /* workspace too small */
#define YCR_E_CUDA (-3)      /* a CUDA call failed; see ycr_last_error() */

/* Anchor grid: level-major, row-major (y outer), centres at (i+0.5)*stride — what
 * make_anchors_polar (utils/tal.py:1393-1407) and Segment.make_anchors (nn/modules/head.py:445-459)
 * produce. */
typedef struct {
    int n_levels;
    int h[YCR_MAX_LEVELS];
    int w[YCR_MAX_LEVELS];
    float stride[YCR_MAX_LEVELS];
} ycr_grid_t;

/* A strided view of the head predictions, one entry per level, so that both layouts on the path are
 * read in place without the reference's cat/permute copies (utils/loss.py:815-822):
 *   raw head feats (B, R+nc, H_l, W_l):  rays = feat_l, cls = feat_l + R*H*W, sb = (R+nc)*H*W,
 *                                         sa = 1, sc = H*W, ray_scale = stride_l, cls_is_logit = 1
 *   assigner API tensors (B,A,R)/(B,A,nc): rays = pd_bboxes + off_l*R, sb = A*R, sa = R, sc = 1,
 *                                         ray_scale = 1, cls_is_logit = 0
 * element (b, a_local, c) of level l lives at ptr[l][b*sb + a_local*sa + c*sc]. */
typedef struct {
    const float* rays[YCR_MAX_LEVELS];
    const float* cls[YCR_MAX_LEVELS];
    int64_t rays_sb[YCR_MAX_LEVELS], rays_sa[YCR_MAX_LEVELS], rays_sc[YCR_MAX_LEVELS];
    int64_t cls_sb[YCR_MAX_LEVELS], cls_sa[YCR_MAX_LEVELS], cls_sc[YCR_MAX_LEVELS];
    float ray_scale[YCR_MAX_LEVELS];
    int cls_is_logit;
} ycr_pred_view_t;

/* Padded ground truth, as v8DetectionLoss.preprocess emits it (utils/loss.py:215-239) and
 * v8SegmentationLoss splits it (utils/loss.py:842-844).  Row (b,g) of each field lives at
 * ptr[(b*G + g) * row_stride]; the split views of the packed (B,G,725) tensor have row_stride 725.
 * mask_gt may be NULL, in which case valid(b,g) = (sum of the 4 box values > 0), utils/loss.py:844.
 * Limits (YCR_E_ARG beyond them, with the numbers in ycr_last_error()): the per-image resolution step keeps
 * 4*A + 244*G bytes (topk = 10) in shared memory, at most 227 KB: G <= 815 GTs per image at A = 8400 (640 px),
 * G <= 401 at A = 33600 (1280 px); G <= 65535 always. */
typedef struct {
    int B, G;
    const float* labels; int64_t labels_stride; /* 1 value: class id as float */
    const float* boxes;  int64_t boxes_stride;  /* 4 values: xyxy px */
    const float* coor;   int64_t coor_stride;   /* 720 values: x0,y0,x1,y1,... px */
    const float* mask_gt; int64_t mask_stride;  /* 1 value or NULL */
} ycr_gt_t;

/* Assigner hyper-parameters, TaskAlignedAssigner.__init__ (utils/tal.py:1125); the live values are
 * topk=10, alpha=0.5, beta=4.0 (utils/loss.py:210). */
typedef struct {
    int topk;        /* 1..64 */
    int num_classes;
    int rays; /* 36 (reference) or 72 */
    float alpha, beta, eps;
} ycr_assign_cfg_t;

/* Dense outputs of TaskAlignedAssigner.forward (utils/tal.py:1204).  Any pointer may be NULL to
 * skip that output.  gt_dist/centerness rows are in (b,g,a) lexicographic order (utils/tal.py:1175)
 * and must have room for B*G*topk rows... see ycr_assign: n_pos is returned in *n_pos_d. */
typedef struct {
    int64_t* target_labels_i64; /* (B,A) */
    float* target_bboxes;       /* (B,A,4) */
    float* target_scores;       /* (B,A,nc) */
    uint8_t* mask_pos;          /* (B,G,A) bool */
    int64_t* target_gt_idx_i64; /* (B,A) */
    uint8_t* fg_mask;           /* (B,A) bool */
    float* gt_dist;             /* (pos_capacity, R) */
    float* centerness;          /* (pos_capacity) */
    int pos_capacity;           /* rows available in gt_dist / centerness */
    int* n_pos_d;               /* device int: number of positives P */
    float* overlaps;            /* optional debug: (B,G,A), get_box_metrics_polar utils/tal.py:1237 */
    float* align_metric;        /* optional debug: (B,G,A) */
} ycr_assign_out_t;

const char* ycr_last_error(void);
int ycr_version(void);
/* sizeof of the seven structs of this header, in declaration order (grid, pred_view, gt, assign_cfg,
 * assign_out, loss_cfg, nms_cfg): lets a foreign-language binding verify its mirror of the layouts.
 * Returns 7. */
int ycr_abi_sizes(int* sizes_out);

/* Measurement hooks (bench.py): while enabled, every kernel of the path is bracketed by CUDA events on
 * the launching stream.  ycr_profile_end waits for them and returns, per kernel tag (YCR_T_* order:
 * 0 gt_setup, 1 cand_overlaps, 2 topk, 3 resolve, 4 positives, 5 loss_stream, 6 finalize, 7 decode,
 * 8 nms_filter, 9 nms_sort, 10 nms_suppress), the summed milliseconds and launch count (16 entries). */
int ycr_profile_begin(int max_records);
/* Restrict the bracketing to the kernel tags whose bit is set (default: all), so that a timed region can
 * carry the events of its dominant kernel only. */
int ycr_profile_select(unsigned tag_mask);
int ycr_profile_end(float* ms_sum_h, int* count_h);
/* Work counters of the candidate kernel since the last reset (synchronises the device); all zero unless the
 * library was built with -DYCR_STATS=1 (measurement build: YCR_NVCC_FLAGS in build.py):
 * out_h[0] candidates swept, [1] (candidate,ray) pairs the own angular bin could not settle,
 * [2] pairs that needed the exact 360-point scan, [3] reserved. */
int ycr_debug_stats(unsigned long long* out_h, int reset);

/* ---- training path --------------------------------------------------------------------------- */

/* Upper bound of in-box candidates (sum over GTs of anchors inside the GT box) computed on the host
 * from host copies of the boxes (B*G rows of xyxy px, row_stride floats apart).  Lets callers size
 * the workspace without a device sync. */
int64_t ycr_candidate_bound_h(const ycr_grid_t* grid, const float* boxes_h, int64_t row_stride, int n_rows);

/* Same bound from the rows the dataloader holds: normalised (x, y, w, h) boxes, utils/loss.py:839. */
int64_t ycr_candidate_bound_xywhn_h(const ycr_grid_t* grid, const float* xywhn_h, int64_t row_stride, int n_rows, float img_w,
                                    float img_h);

size_t ycr_assign_workspace_bytes(const ycr_grid_t* grid, int B, int G, const ycr_assign_cfg_t* cfg,
                                  int64_t cand_capacity);

/* Replaces TaskAlignedAssigner.forward (utils/tal.py:1135-1204): in-box candidates
 * (utils/tal.py:52-66), polygon->polar targets + Polar-IoU per candidate (utils/tal.py:1237-1284,
 * 1445-1464), per-GT top-k (utils/tal.py:1304-1338), multi-GT resolution (utils/tal.py:214-248),
 * polar targets of the positives (utils/tal.py:1172-1193), targets and normalisation
 * (utils/tal.py:1340-1390, 1197-1202). */
int ycr_assign(const ycr_grid_t* grid, const ycr_pred_view_t* pred, const ycr_gt_t* gt,
               const ycr_assign_cfg_t* cfg, const ycr_assign_out_t* out, void* workspace,
               size_t workspace_bytes, int64_t cand_capacity, void* stream);

/* Hyper-parameters of v8SegmentationLoss (utils/loss.py:876-877, cfg/default.yaml:89-90). */
typedef struct {
    float box_gain; /* hyp.box = 7.5 */
    float cls_gain; /* hyp.cls = 0.5 */
} ycr_loss_cfg_t;

size_t ycr_seg_loss_workspace_bytes(const ycr_grid_t* grid, int B, int G, const ycr_assign_cfg_t* cfg,
                                    int64_t cand_capacity);

/* Replaces v8SegmentationLoss.__call__ (utils/loss.py:808-878) from the head feature maps on:
 * assignment as above, BCE-with-logits class loss (utils/loss.py:866-867), Polar-IoU log-ratio loss
 * (MaskIOULoss.forward utils/loss.py:113-127), gains, and — in the same pass — the gradient of the
 * returned scalar `loss.sum()*B` with respect to every feature map element.
 *   feats[l]      (B, R+nc, H_l, W_l) contiguous
 *   grad_feats[l] same shape, fully overwritten (may be NULL for a forward-only call)
 *   loss_out      device float[4]: {total = (l_box+l_cls)*B, l_box, l_cls, target_scores_sum} */
int ycr_seg_loss_fwd_bwd(const ycr_grid_t* grid, const float* const* feats, float* const* grad_feats,
                         const ycr_gt_t* gt, const ycr_assign_cfg_t* acfg, const ycr_loss_cfg_t* lcfg,
                         float* loss_out, void* workspace, size_t workspace_bytes, int64_t cand_capacity,
                         void* stream);

/* Same for feature maps of element type `dtype` (0 fp32, 1 fp16, 2 bf16) - what the head emits under autocast, the
 * reference's default (engine/trainer.py:332).  The maps are read in place, arithmetic is fp32 as in the reference
 * (its sigmoid and its BCE targets are rounded to the input type, utils/loss.py:861,867 - so are they here), and
 * grad_feats are written in the input type. */
int ycr_seg_loss_fwd_bwd_dt(const ycr_grid_t* grid, const void* const* feats, void* const* grad_feats, int dtype,
                            const ycr_gt_t* gt, const ycr_assign_cfg_t* acfg, const ycr_loss_cfg_t* lcfg, float* loss_out,
                            void* workspace, size_t workspace_bytes, int64_t cand_capacity, void* stream);
int ycr_scale_grads_dt(const ycr_grid_t* grid, int B, int channels, void* const* grad_feats, int dtype, const float* scale_d,
                       void* stream);

/* In-place grad *= *scale_d for the three gradient maps; returns immediately on the device when
 * *scale_d == 1 (the common loss.backward() case), so autograd's upstream gradient costs no pass. */
int ycr_scale_grads(const ycr_grid_t* grid, int B, int channels, float* const* grad_feats,
                    const float* scale_d, void* stream);

/* GT packing, replaces v8DetectionLoss.preprocess (utils/loss.py:215-239): `targets` rows exactly as
 * utils/loss.py:839 concatenates them — [image index, class, x, y, w, h (normalised), 720 contour values
 * (normalised)], row_stride floats apart, any order of images — -> padded (B,G,725) in px. */
int ycr_pack_targets(const float* targets, int64_t row_stride, int N, int B, int G, float img_w, float img_h,
                     float* out_packed, void* stream);

/* Same with the two parts of a row in separate arrays - [image index, class, x, y, w, h] (head_stride floats apart) and
 * the 720 contour values (seg_stride floats apart) - so that a caller can stage `batch['segments']` with one copy. */
int ycr_pack_targets_split(const float* head, int64_t head_stride, const float* segments, int64_t seg_stride, int N, int B, int G,
                           float img_w, float img_h, float* out_packed, void* stream);

/* Host-side half of GT packing (everything in host memory, no CUDA call): lays the rows of a collated batch out in
 * one staging buffer - N x 6 header floats [image index, class, x, y, w, h] followed by N x 720 contour floats, the
 * layout ycr_pack_targets_split takes after one host-to-device copy - and returns G = the largest number of boxes
 * of one image (utils/loss.py:224-226) and the candidate bound.  batch['segments'] arrives as a list of per-image
 * tensors: seg_ptrs_h[k] holds seg_rows_h[k] rows of 720 floats.  Behind the rows it writes the int32 table
 * row_of[b*G + slot] (row index or -1 for padding) that ycr_pack_targets_mapped takes.  staging_h needs
 * N * 726 floats + B * G ints, G <= N (pinned, so that the copy is asynchronous). */
int ycr_stage_targets_h(const float* batch_idx_h, const float* cls_h, const float* bboxes_h, const float* const* seg_ptrs_h,
                        const int* seg_rows_h, int n_seg_tensors, int N, int B, const ycr_grid_t* grid, float img_w, float img_h,
                        float* staging_h, int* G_out, int64_t* cand_bound_out);

/* GT packing from staged rows and the row table of ycr_stage_targets_h: one block per padded (image, slot) row, every
 * element of out_packed (B, G, 725) is written (no memset). */
int ycr_pack_targets_mapped(const float* head, int64_t head_stride, const float* segments, int64_t seg_stride, const int* row_of,
                            int B, int G, float img_w, float img_h, float* out_packed, void* stream);

/* Contour resampling, replaces ops.resample_segments (utils/ops.py:676-693; n = 360 at
 * utils/instance.py:202): S open polygons stored back to back in pts (total,2); polygon s owns rows
 * offsets[s] .. offsets[s+1]-1.  Each is closed and linearly resampled to n_out points ->
 * out (S, n_out, 2) float32, bit-identical to numpy's double-precision np.interp result. */
int ycr_resample_segments(const float* pts, const int* offsets, int S, int n_out, float* out, void* stream);

/* Box terms kept for API coverage (dormant on the live polar path): replaces BboxLoss.forward
 * (utils/loss.py:61-75) with _df_loss (:77-87), bbox_iou(CIoU) (utils/metrics.py:77-130) and bbox2dist
 * (utils/tal.py:1437-1440), forward and gradient in one pass.
 *   pred_dist (B,A,4*(reg_max+1)) logits, pred_bboxes/target_bboxes (B,A,4) xyxy, anchor_points (A,2),
 *   target_scores (B,A,nc), fg_mask (B,A) bool, target_scores_sum_d device scalar.
 *   loss_out device float[2] = {loss_iou, loss_dfl}; grad_* same shapes as the inputs, fully overwritten
 *   (d(loss_iou + loss_dfl)/d input), may be NULL. */
size_t ycr_bbox_loss_workspace_bytes(int B, int A);
int ycr_bbox_loss_fwd_bwd(const float* pred_dist, const float* pred_bboxes, const float* anchor_points,
                          const float* target_bboxes, const float* target_scores, const uint8_t* fg_mask,
                          const float* target_scores_sum_d, int B, int A, int nc, int reg_max, int use_dfl,
                          float* loss_out, float* grad_pred_dist, float* grad_pred_bboxes, void* workspace,
                          size_t workspace_bytes, void* stream);

/* ---- inference path -------------------------------------------------------------------------- */

/* Replaces Segment.forward eval branch / distance2mask (nn/modules/head.py:461-494, 559-570):
 * feats -> allpred (B, 4+nc+3R, A) = [box xyxy | sigmoid cls | x_0.. | y_0.. | valid_0..]. */
int ycr_decode(const ycr_grid_t* grid, const float* const* feats, int B, int nc, int R, float* allpred,
               void* stream);
/* Same, and also writes per anchor the best class as {float score, int class} pairs (B, A, 8 bytes; first
 * maximum, i.e. `cls.max(1)` of utils/ops.py:386; class -1 = not provided).  Passed to ycr_nms through
 * ycr_nms_cfg_t.best_class it saves single-label NMS the read of all class rows. */
int ycr_decode_best(const ycr_grid_t* grid, const float* const* feats, int B, int nc, int R, float* allpred,
                    void* best_class_out, void* stream);

/* Same for feature maps of element type `dtype` (0 fp32, 1 fp16, 2 bf16: the validator runs the model in half,
 * engine/validator.py:103-104); arithmetic and allpred stay fp32.  best_class_out may be NULL. */
int ycr_decode_dt(const ycr_grid_t* grid, const void* const* feats, int dtype, int B, int nc, int R, float* allpred,
                  void* best_class_out, void* stream);

typedef struct {
    float conf_thres, iou_thres;
    int agnostic, multi_label;
    int max_det, nc, max_nms; /* 1 <= max_det <= 1024 (YCR_E_ARG beyond: the kept list lives in shared memory) */
    float max_wh;
    const int* classes; int n_classes; /* optional device list of class ids to keep, utils/ops.py:390 */
    int compact_rows;   /* 0: image b's rows start at out_rows[b*max_det]; 1: rows of all images back to back
                         * (image b starts at row sum(out_counts[0..b-1])), so the host can split one tensor */
    const void* best_class; /* optional: what ycr_decode_best wrote for THIS prediction tensor, else NULL */
    /* optional: the head feature maps THIS prediction tensor was decoded from (ycr_decode*, same grid, element type
     * feats_dtype) - the kept rows are then recomputed from the R ray values of their anchor instead of being
     * collected element by element from the channel-major prediction (a sector per element); NULL entries = off */
    const void* feats[YCR_MAX_LEVELS];
    const ycr_grid_t* grid;
    int feats_dtype;
    int rays;
} ycr_nms_cfg_t;

size_t ycr_nms_workspace_bytes(int B, int A, int channels, const ycr_nms_cfg_t* cfg);

/* Replaces ops.non_max_suppression, polar variant (utils/ops.py:285-424) including the
 * torchvision.ops.nms step (utils/ops.py:407).  prediction (B, 4+nc+nm, A).
 *   out_rows   B*max_det rows of 6+nm floats: kept rows in descending-score order, image b's rows at
 *              row b*max_det (cfg->compact_rows == 0) or directly behind image b-1's (compact_rows == 1)
 *   out_counts (B) device int: rows kept per image */
int ycr_nms(const float* prediction, int B, int channels, int A, const ycr_nms_cfg_t* cfg, float* out_rows,
            int* out_counts, void* workspace, size_t workspace_bytes, void* stream);

/* ---- deployment: head feature maps -> detections in one call -------------------------------- */

/* What an inference plugin (TensorRT IPluginV2 enqueue, an ONNX Runtime custom op, the C++ demo of the reference's
 * examples/) calls behind the export form of the head (nn/modules/head.py:572-574 returns the raw ray and class
 * maps): ycr_decode + ycr_nms (single-label, as models/yolo/segment/predict.py:18 calls it) without ever writing
 * the (B, 4+nc+3R, A) prediction tensor - best class per anchor from the class logits, conf filter and sort, boxes
 * from the rays of the candidates only, suppression, and the kept rows [box | conf | class | x_0.. | y_0.. | valid_0..]
 * from the rays of the kept anchors.  Rows and counts are identical to the two-call form.  cfg->nc, best_class,
 * feats, grid are ignored (taken from the arguments); multi_label is rejected. */
size_t ycr_detect_workspace_bytes(const ycr_grid_t* grid, int B, const ycr_nms_cfg_t* cfg);
int ycr_detect(const ycr_grid_t* grid, const void* const* feats, int dtype, int B, int nc, int R, const ycr_nms_cfg_t* cfg,
               float* out_rows, int* out_counts, void* workspace, size_t workspace_bytes, void* stream);

/* ---- validation: contour -> mask rasterisation and mask IoU ------------------------------ */

/* Replaces the fill loop ops.process_mask has commented out (utils/ops.py:768-825, :794-809): rows are NMS output
 * rows [box4 | conf | cls | x_0.. | y_0.. | valid_0..] (row_stride floats apart, 6+3R used); for each of the n
 * detections the valid contour points are truncated to int32 and filled as cv2.fillPoly does (boundary lines +
 * even-odd scan lines, 16.16 fixed point) into masks (n, H, W) uint8, values 0/1. */
int ycr_rasterize_contours(const float* rows, int64_t row_stride, int n, int R, int H, int W, uint8_t* masks, void* stream);

/* Replaces metrics.mask_iou (utils/metrics.py:133-155): mask1 (N, n) and mask2 (M, n), uint8 (dtype 0) or
 * float32 (dtype 1), non-zero = set -> iou (N, M) = inter / (area1 + area2 - inter + eps) in fp32.  Both sets are
 * bit-packed into the workspace, intersections are popcounts. */
size_t ycr_mask_iou_workspace_bytes(int N, int M, int64_t n);
int ycr_mask_iou(const void* mask1, int dtype1, const void* mask2, int dtype2, int N, int M, int64_t n, float eps, float* iou,
                 void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* YCR_B200_H */
