// Internal interface between the C-ABI wrappers (api.cu) and the training-path kernels.
#pragma once
#include "common.cuh"
#include "polar_core.cuh"
#include "dtype.cuh"

// Everything the assignment stage leaves in the workspace.
struct AssignWs {
    // K0: per-GT candidate rectangles
    int4* rect;        // [BG][L] (x0, y0, w, h) in grid cells
    int* ncand;        // [BG]
    int* cand_off;     // [BG+1]
    int* chunk_off;    // [BG+1]
    int* chunk_bg;     // [chunks_cap] hand-out order of K1: GT of the u-th chunk drawn ...
    int* chunk_work;   // [chunks_cap] ... and its chunk index (full chunks first, the cheaper partial ones last)
    int* part_off;     // [BG+1] number of GTs before this one that have a partial chunk
    int chunks_cap;
    uint8_t* valid;    // [BG]
    int* totals;       // [4] = {M, T, K1 work counter, -}
    int* err;          // [1] device error flag (candidate capacity exceeded)
    // K1: per-candidate metrics
    float* cand_align; // [cap]
    float* cand_ov;    // [cap]
    float* cand_t;     // [chunks][R][K1 block] ray targets of every candidate, or null (then positives are re-swept)
    // K2: per-GT top-k
    int* sel;          // [BG][topk] anchor index or -1
    // K3: per-image positives, padded to pos_cap = G*topk rows per image
    int* npos;         // [B]
    int* pos_anchor;   // [B][pos_cap]
    int* pos_g;        // [B][pos_cap]
    float* pos_norm;   // [B][pos_cap]  normalised target score
    int* pos_row;      // [B][A] row of the positive at this anchor or -1
    int* gt_row_start; // [BG]
    int* gt_row_cnt;   // [BG]
    float* tss_part;   // [B] per-image sum of target scores
    int* img_base;     // [B+1] exclusive scan of npos
    float* tss;        // [2] {max(sum,1), sum}
    // K4: per-positive loss pieces
    float* pos_loss;   // [B][pos_cap]
    float* pos_grad;   // [B][pos_cap][R]  d(total)/d(raw ray)
    // K5
    float* bce_part;   // [n_bce_blocks]
    int pos_cap;
    int64_t cand_cap;
    int n_bce_blocks;
};

size_t assign_ws_layout(AssignWs* ws, void* base, const GridDev& grid, int B, int G, int topk, int R,
                        int64_t cand_cap, bool with_loss);

struct AssignArgs {
    GridDev grid;
    ycr_pred_view_t pred;
    ycr_gt_t gt;
    ycr_assign_cfg_t cfg;
    PolarConst pc;
    int dtype;   // YCR_F32 / YCR_F16 / YCR_BF16: element type behind pred.rays / pred.cls (fused loss: of the feature maps)
};

int launch_assign_core(const AssignArgs& a, const AssignWs& ws, int* n_pos_d, cudaStream_t st);
int launch_assign_dense(const AssignArgs& a, const AssignWs& ws, const ycr_assign_out_t& out, cudaStream_t st);
int launch_positive_targets(const AssignArgs& a, const AssignWs& ws, float* gt_dist, float* centerness,
                            int pos_capacity, bool with_loss, const ycr_loss_cfg_t* lcfg,
                            cudaStream_t st);
int launch_loss_stream(const AssignArgs& a, const AssignWs& ws, const void* const* feats, void* const* grads,
                       const ycr_loss_cfg_t& lcfg, float* loss_out, cudaStream_t st);
