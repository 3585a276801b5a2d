"""CPU ORACLE for the polar-contour hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import this module; the product path (`yolo-contour-regression_b200/`) never does and fails loudly
when its CUDA library is missing.

What it is: a from-scratch restatement, in torch-CPU fp32 ops, of the algorithm of the reference
fork ai4in/YOLO-Contour-Regression (paths below are relative to
`/root/reference/ultralytics-main/ultralytics/`).  torch-CPU (rather than numpy) is used on purpose:
the reference *is* torch fp32, so the same primitive ops (atan2, topk, sum order) give the closest
possible numerics, and the restatement doubles as the multi-threaded CPU baseline.

Differences from the reference, all deliberate:
  * the number of rays R is a parameter (the reference hard-codes 36 at utils/tal.py:1178,1263);
  * work is chunked (by image and by candidate block) so memory stays bounded — the reference
    materialises an (M,R,360) temporary and OOMs beyond B~32;
  * ties are broken lowest-index (stable sorts) where torch.topk leaves them unspecified;
  * an all-empty batch returns a consistent 8-tuple instead of the reference's 6-tuple
    (utils/tal.py:1157-1161, which its only caller cannot unpack);
  * every selection step also reports a *tie margin* so tests can restrict bit-exact claims to
    inputs where the reference's own result is well defined ("tie-free synthetic data").

Pinning: the reference ships no golden vectors or unit tests for this path, so
parity is pinned against OUTPUTS OF THE REFERENCE ITSELF, generated in the build container by
`tests/golden/make_golden.py` (imports /root/reference) and committed under `tests/golden/`;
`tests/test_oracle_golden.py` checks this oracle against every one of them.
NMS: the suppression step lives in torchvision (`torchvision.ops.nms`, requirement
`torchvision>=0.8.1`, installed 0.26.0) and is restated in `greedy_nms`; it is cross-checked
against the installed torchvision CPU kernel in the CPU tests.
"""
from __future__ import annotations

import math

import numpy as np
import torch

CONTOUR_POINTS = 360
NEAREST_K = 4          # utils/tal.py:1183,1268  (k=4)
GATE_DEG = 3.0         # utils/tal.py:1185,1270  (.gt(3))
FLOOR = 1e-6           # utils/tal.py:1189-1191, 1275-1277, 1455


# --------------------------------------------------------------------------------------------
# anchors and GT packing
# --------------------------------------------------------------------------------------------
def make_anchors(level_shapes, strides, offset: float = 0.5):
    """utils/tal.py:1393-1407 (make_anchors_polar) and nn/modules/head.py:445-459.
    Returns anchor centres in grid units (A,2) [x,y], strides (A,1), level-major, row-major."""
    pts, st = [], []
    for (h, w), s in zip(level_shapes, strides):
        sx = torch.arange(w, dtype=torch.float32) + offset
        sy = torch.arange(h, dtype=torch.float32) + offset
        yy, xx = torch.meshgrid(sy, sx, indexing="ij")
        pts.append(torch.stack((xx, yy), -1).view(-1, 2))
        st.append(torch.full((h * w, 1), float(s), dtype=torch.float32))
    return torch.cat(pts), torch.cat(st)


def pack_targets(batch: dict, batch_size: int, img_hw):
    """utils/loss.py:834-844 + v8DetectionLoss.preprocess utils/loss.py:215-239.
    -> (B, G, 5+720): [cls | xyxy px | 720 contour values px].  The contour columns are scaled as
    the reference does it: columns 5:365 by image width and 365:725 by image height, although the
    data is x,y-interleaved (harmless for square images; restated literally)."""
    h, w = float(img_hw[0]), float(img_hw[1])
    n = batch["batch_idx"].numel()
    width = 5 + 2 * CONTOUR_POINTS
    if n == 0:
        return torch.zeros(batch_size, 0, width)
    seg = torch.cat([s for s in batch["segments"]]).contiguous().view(-1, 2 * CONTOUR_POINTS)
    t = torch.cat((batch["batch_idx"].view(-1, 1), batch["cls"].view(-1, 1),
                   batch["bboxes"], seg), 1).float()
    idx = t[:, 0].long()
    counts = torch.bincount(idx, minlength=batch_size)
    out = torch.zeros(batch_size, int(counts.max()), width)
    for j in range(batch_size):
        m = idx == j
        k = int(m.sum())
        if k:
            out[j, :k] = t[m, 1:]
    xywh = out[..., 1:5] * torch.tensor([w, h, w, h])
    xy, wh = xywh[..., :2], xywh[..., 2:] / 2
    out[..., 1:5] = torch.cat((xy - wh, xy + wh), -1)  # utils/ops.py xywh2xyxy
    out[..., 5:5 + CONTOUR_POINTS] = out[..., 5:5 + CONTOUR_POINTS] * w
    out[..., 5 + CONTOUR_POINTS:] = out[..., 5 + CONTOUR_POINTS:] * h
    return out


def in_box_mask(anc_px, gt_boxes, eps: float = 1e-9):
    """utils/tal.py:52-66 select_candidates_in_gts.  anc_px (A,2), gt_boxes (B,G,4) -> bool (B,G,A)"""
    lt, rb = gt_boxes[..., None, :2], gt_boxes[..., None, 2:]
    d = torch.cat((anc_px[None, None] - lt, rb - anc_px[None, None]), -1)
    return d.amin(-1) > eps


# --------------------------------------------------------------------------------------------
# polygon -> polar targets
# --------------------------------------------------------------------------------------------
def ray_angles_deg(R: int):
    """utils/tal.py:1178: theta = arange(0, 360, 360//36)"""
    return torch.arange(0, 360, 360 // R, dtype=torch.float32)


def polar_targets(anc_px, contour, R: int = 36, block: int = 2048, tol_deg: float = 2e-4):
    """utils/tal.py:1257-1277 (all candidates) == utils/tal.py:1172-1193 (positives), get_angle
    utils/tal.py:1286-1301.

    anc_px (M,2) anchor centres in px; contour (M,360,2) px.
    Returns dict: t (M,R) ray targets; t_lo/t_hi (M,R) the envelope of results reachable when a
    selection whose margin is below `tol_deg` flips (== t where the result is unambiguous);
    ambiguous (M,R) bool = t_lo != t_hi.
    """
    M = anc_px.shape[0]
    theta = ray_angles_deg(R)
    t_out = torch.empty(M, R)
    lo_out = torch.empty(M, R)
    hi_out = torch.empty(M, R)
    for s in range(0, M, block):
        a = anc_px[s:s + block, None, :]                       # (m,1,2)
        c = contour[s:s + block]                               # (m,360,2)
        v = c - a
        ang = torch.atan2(v[..., 1], v[..., 0])
        ang = ang * 180.0 / np.pi                              # op order as utils/tal.py:1297
        ang = torch.where(ang < 0, ang + 360, ang)
        diff = (ang[:, None, :] - theta[None, :, None]).abs()  # (m,R,360)
        diff = torch.where(diff > 180.0, 360 - diff, diff)
        val, idx = torch.topk(diff, NEAREST_K + 1, dim=2, largest=False)  # sorted ascending
        del diff
        dist = torch.norm(v, 2, 2)                             # (m,360)
        d5 = torch.gather(dist[:, None, :].expand(-1, R, -1), 2, idx)     # (m,R,5)
        gated = val[..., 0] > GATE_DEG                          # min over the 4 == first of sorted
        floor = torch.full_like(d5[..., 0], FLOOR)
        t_sel = torch.where(gated, floor, d5[..., :NEAREST_K].amax(2)).clamp(min=FLOOR)
        # alternative outcomes under small angular perturbations
        t_swap = torch.where(gated, floor, torch.maximum(d5[..., :NEAREST_K - 1].amax(2),
                                                         d5[..., NEAREST_K])).clamp(min=FLOOR)
        near_swap = (val[..., NEAREST_K] - val[..., NEAREST_K - 1]) < tol_deg
        t_gateflip = torch.where(gated, d5[..., :NEAREST_K].amax(2).clamp(min=FLOOR), floor)
        near_gate = (val[..., 0] - GATE_DEG).abs() < tol_deg
        lo = t_sel.clone()
        hi = t_sel.clone()
        lo = torch.where(near_swap, torch.minimum(lo, t_swap), lo)
        hi = torch.where(near_swap, torch.maximum(hi, t_swap), hi)
        lo = torch.where(near_gate, torch.minimum(lo, t_gateflip), lo)
        hi = torch.where(near_gate, torch.maximum(hi, t_gateflip), hi)
        # earlier near-ties inside the first four only permute the selected set -> same max
        t_out[s:s + block] = t_sel
        lo_out[s:s + block] = lo
        hi_out[s:s + block] = hi
    return {"t": t_out, "t_lo": lo_out, "t_hi": hi_out, "ambiguous": lo_out != hi_out}


def polar_iou(target, pred):
    """utils/tal.py:1445-1464 MaskIOU(target, pred): sum(clamp(min,1e-6)) / sum(max) over rays."""
    both = torch.stack([pred, target], -1)
    l_max = both.max(dim=-1)[0]
    l_min = both.min(dim=-1)[0].clamp(min=FLOOR)
    return l_min.sum(dim=-1) / l_max.sum(dim=-1)


def centerness(t):
    """utils/tal.py:1220-1226 polar_centerness_target"""
    return torch.sqrt(t.min(dim=-1)[0] / t.max(dim=-1)[0])


# --------------------------------------------------------------------------------------------
# TaskAlignedAssigner.forward
# --------------------------------------------------------------------------------------------
def assign(pd_scores, pd_rays, anc_px, gt_labels, gt_boxes, mask_gt, gt_coor,
           topk: int = 10, alpha: float = 0.5, beta: float = 4.0, eps: float = 1e-9,
           R: int | None = None, rel_tol: float = 2e-5, tol_deg: float = 2e-4):
    """utils/tal.py:1135-1204 (forward), :1206-1218 (get_pos_mask), :1237-1284
    (get_box_metrics_polar), :1304-1338 (select_topk_candidates), :214-248
    (select_highest_overlaps), :1340-1390 (get_targets), normalisation :1197-1202.

    pd_scores (B,A,nc) sigmoid scores; pd_rays (B,A,R) px; anc_px (A,2) px; gt_labels (B,G,1);
    gt_boxes (B,G,4) xyxy px; mask_gt (B,G,1); gt_coor (B,G,720) px interleaved x,y.
    Images are processed one at a time (values are image-independent).

    Returns a dict with the reference's 8 outputs plus intermediates and the tie-margin report:
      certain (B,) bool — every discrete decision of that image is stable under perturbations of
      the polar targets within their ambiguity envelope and `rel_tol` relative noise.
    """
    B, A, nc = pd_scores.shape
    G = gt_boxes.shape[1]
    R = pd_rays.shape[-1] if R is None else R
    out = {
        "target_labels": torch.zeros(B, A, dtype=torch.int64),
        "target_bboxes": torch.zeros(B, A, 4),
        "target_scores": torch.zeros(B, A, nc),
        "mask_pos": torch.zeros(B, G, A, dtype=torch.bool),
        "target_gt_idx": torch.zeros(B, A, dtype=torch.int64),
        "fg_mask": torch.zeros(B, A, dtype=torch.bool),
        "overlaps": torch.zeros(B, G, A),
        "overlaps_lo": torch.zeros(B, G, A),
        "overlaps_hi": torch.zeros(B, G, A),
        "align_metric": torch.zeros(B, G, A),
        "certain": torch.ones(B, dtype=torch.bool),
        "n_candidates": 0, "n_ambiguous_rays": 0,
    }
    dist_rows, cent_rows, amb_rows = [], [], []
    if G == 0:
        out["target_labels"].fill_(nc)  # bg_idx, utils/tal.py:1159
        out["gt_dist"] = torch.zeros(0, R)
        out["centerness"] = torch.zeros(0)
        out["gt_dist_ambiguous"] = torch.zeros(0, R, dtype=torch.bool)
        return out
    for b in range(B):
        boxes = gt_boxes[b]                                    # (G,4)
        valid = mask_gt[b, :, 0].bool()                        # (G,)
        contour = gt_coor[b].view(G, CONTOUR_POINTS, 2)
        in_gts = in_box_mask(anc_px, boxes[None])[0]           # (G,A)
        cand = in_gts & valid[:, None]
        gi, ai = torch.nonzero(cand, as_tuple=True)            # (g,a) lexicographic
        out["n_candidates"] += int(gi.numel())
        overlaps = torch.zeros(G, A)
        ov_lo = torch.zeros(G, A)
        ov_hi = torch.zeros(G, A)
        scores = torch.zeros(G, A)
        if gi.numel():
            pt = polar_targets(anc_px[ai], contour[gi], R, tol_deg=tol_deg)
            pr = pd_rays[b, ai]
            overlaps[gi, ai] = polar_iou(pt["t"], pr)
            lo_num = torch.minimum(pr, pt["t_lo"]).clamp(min=FLOOR).sum(-1)
            hi_num = torch.minimum(pr, pt["t_hi"]).clamp(min=FLOOR).sum(-1)
            lo_den = torch.maximum(pr, pt["t_lo"]).sum(-1)
            hi_den = torch.maximum(pr, pt["t_hi"]).sum(-1)
            ov_lo[gi, ai] = lo_num / hi_den
            ov_hi[gi, ai] = hi_num / lo_den
            scores[gi, ai] = pd_scores[b, ai, gt_labels[b, gi, 0].long()]
            out["n_ambiguous_rays"] += int(pt["ambiguous"].sum())
        align = scores.pow(alpha) * overlaps.pow(beta)
        al_lo = scores.pow(alpha) * ov_lo.pow(beta) * (1 - rel_tol)
        al_hi = scores.pow(alpha) * ov_hi.pow(beta) * (1 + rel_tol)
        # --- per-GT top-k over anchors, lowest index on ties (stable sort) ---
        order = torch.sort(align, dim=1, descending=True, stable=True)[1][:, :topk]   # (G,topk)
        sel = torch.zeros(G, A, dtype=torch.bool)
        sel[torch.arange(G)[:, None].expand(-1, order.shape[1]), order] = True
        sel &= valid[:, None]            # invalid rows: idx->0, count>1 -> dropped (tal.py:1325,1336)
        certain = True
        for g in torch.nonzero(valid).flatten().tolist():
            s = sel[g]
            inside = s & in_gts[g]
            rest = (~s) & in_gts[g]
            if inside.any() and rest.any():
                if al_lo[g][inside].min() <= al_hi[g][rest].max():
                    certain = False
            # a zero-metric in-box anchor picked (or not) as filler depends on topk's tie order
            if (s & in_gts[g] & (align[g] == 0)).any() or \
               (int((align[g] > 0).sum()) < topk and (rest & (align[g] == 0)).any()):
                certain = False
        mask_pos = (sel & in_gts & valid[:, None]).float()
        # --- select_highest_overlaps ---
        fg = mask_pos.sum(0)
        if fg.max() > 1:
            multi = (fg[None] > 1).expand(G, -1)
            best = overlaps.argmax(0)
            one_hot = torch.zeros(G, A)
            one_hot.scatter_(0, best[None], 1)
            mask_pos = torch.where(multi, one_hot, mask_pos).float()
            cols = torch.nonzero(fg > 1).flatten()
            bi = overlaps[:, cols].argmax(0)
            lo_best = ov_lo[bi, cols]
            hi_others = ov_hi[:, cols].clone()
            hi_others[bi, torch.arange(cols.numel())] = -1
            if (lo_best * (1 - rel_tol) <= hi_others.max(0)[0] * (1 + rel_tol)).any():
                certain = False
            fg = mask_pos.sum(0)
        tgi = mask_pos.argmax(0)
        # --- polar targets of the positives, (g,a) order ---
        pg, pa = torch.nonzero(mask_pos.bool(), as_tuple=True)
        if pg.numel():
            ptp = polar_targets(anc_px[pa], contour[pg], R, tol_deg=tol_deg)
            dist_rows.append(ptp["t"])
            cent_rows.append(centerness(ptp["t"]))
            amb_rows.append(ptp["ambiguous"])
        # --- get_targets ---
        labels = gt_labels[b, :, 0].long()[tgi].clamp(min=0)
        tboxes = boxes[tgi]
        tscores = torch.zeros(A, nc)
        tscores.scatter_(1, labels[:, None], 1.0)
        tscores = torch.where((fg > 0)[:, None], tscores, torch.zeros(()))
        # --- normalisation ---
        al = align * mask_pos
        pos_al = al.amax(1, keepdim=True)
        pos_ov = (overlaps * mask_pos).amax(1, keepdim=True)
        norm = (al * pos_ov / (pos_al + eps)).amax(0)
        tscores = tscores * norm[:, None]
        out["target_labels"][b] = labels
        out["target_bboxes"][b] = tboxes
        out["target_scores"][b] = tscores
        out["mask_pos"][b] = mask_pos.bool()
        out["target_gt_idx"][b] = tgi
        out["fg_mask"][b] = fg > 0
        out["overlaps"][b] = overlaps
        out["overlaps_lo"][b] = ov_lo
        out["overlaps_hi"][b] = ov_hi
        out["align_metric"][b] = align
        out["certain"][b] = certain
    out["gt_dist"] = torch.cat(dist_rows) if dist_rows else torch.zeros(0, R)
    out["centerness"] = torch.cat(cent_rows) if cent_rows else torch.zeros(0)
    out["gt_dist_ambiguous"] = torch.cat(amb_rows) if amb_rows else torch.zeros(0, R, dtype=torch.bool)
    return out


# --------------------------------------------------------------------------------------------
# v8SegmentationLoss.__call__
# --------------------------------------------------------------------------------------------
def polar_iou_loss(pred_rays, target_rays, target_scores, target_scores_sum):
    """utils/loss.py:113-127 MaskIOULoss.forward"""
    weight = target_scores.sum(-1)
    both = torch.stack([pred_rays, target_rays], -1)
    l_max = both.max(dim=2)[0]
    l_min = both.min(dim=2)[0].clamp(min=FLOOR)
    loss = (l_max.sum(dim=1) / l_min.sum(dim=1)).log() * weight
    return loss.sum() / target_scores_sum


def seg_loss(feats, batch, strides=(8, 16, 32), nc: int = 80, R: int = 36,
             box_gain: float = 7.5, cls_gain: float = 0.5, topk: int = 10,
             alpha: float = 0.5, beta: float = 4.0, with_grad: bool = True):
    """utils/loss.py:808-878 v8SegmentationLoss.__call__ (+ autograd for the gradients).

    feats: list of (B, R+nc, H_l, W_l).  Returns dict(loss, loss_items (2,), grads list, assign)."""
    feats = [f.detach().clone().requires_grad_(with_grad) for f in feats]
    B = feats[0].shape[0]
    no = R + nc
    cat = torch.cat([f.view(B, no, -1) for f in feats], 2)
    pred_rays, pred_logits = cat.split((R, nc), 1)
    pred_logits = pred_logits.permute(0, 2, 1).contiguous()
    pred_rays = pred_rays.permute(0, 2, 1).contiguous()
    img_hw = (feats[0].shape[2] * strides[0], feats[0].shape[3] * strides[0])
    level_shapes = [tuple(f.shape[2:]) for f in feats]
    anc, st = make_anchors(level_shapes, strides)
    targets = pack_targets(batch, B, img_hw)
    gt_labels, gt_boxes, gt_coor = targets.split((1, 4, 2 * CONTOUR_POINTS), 2)
    mask_gt = (gt_boxes.sum(2, keepdim=True) > 0).float()
    pred_px = pred_rays * st
    loss = torch.zeros(2)
    if targets.shape[1] == 0:
        asg = assign(pred_logits.detach().sigmoid(), pred_px.detach(), anc * st, gt_labels,
                     gt_boxes, mask_gt, gt_coor, topk, alpha, beta, R=R)
        tss = torch.tensor(1.0)
        loss[1] = torch.nn.functional.binary_cross_entropy_with_logits(
            pred_logits, asg["target_scores"], reduction="none").sum() / tss
    else:
        asg = assign(pred_logits.detach().sigmoid(), pred_px.detach(), anc * st, gt_labels,
                     gt_boxes, mask_gt, gt_coor, topk, alpha, beta, R=R)
        tscores = asg["target_scores"]
        tss = max(tscores.sum(), 1)
        loss[1] = torch.nn.functional.binary_cross_entropy_with_logits(
            pred_logits, tscores, reduction="none").sum() / tss
        mp = asg["mask_pos"]
        if mp.sum():
            G = mp.shape[1]
            pr = pred_px.unsqueeze(1).expand(-1, G, -1, -1)[mp]
            ts = tscores.unsqueeze(1).expand(-1, G, -1, -1)[mp]
            loss[0] = polar_iou_loss(pr, asg["gt_dist"], ts, tss)
    loss[0] = loss[0] * box_gain
    loss[1] = loss[1] * cls_gain
    total = loss.sum() * B
    grads = None
    if with_grad:
        total.backward()
        grads = [f.grad for f in feats]
    return {"loss": total.detach(), "loss_items": loss.detach(), "grads": grads, "assign": asg,
            "target_scores_sum": float(tss)}


# --------------------------------------------------------------------------------------------
# Segment decode and NMS
# --------------------------------------------------------------------------------------------
def decode(feats, strides=(8, 16, 32), nc: int = 80, R: int = 36):
    """nn/modules/head.py:461-494 distance2mask (+ make_anchors :445-459, forward eval :559-570).
    -> (B, 4+nc+3R, A): [box xyxy | sigmoid cls | x_0..x_{R-1} | y_0..y_{R-1} | valid_0..]"""
    B = feats[0].shape[0]
    no = R + nc
    level_shapes = [tuple(f.shape[2:]) for f in feats]
    anc, st = make_anchors(level_shapes, strides)
    pts = anc * st
    d = torch.cat([f.view(B, no, -1) for f in feats], 2).permute(0, 2, 1)
    rays, cls = d.split((R, nc), -1)
    ang = torch.arange(0, 360, 360 // R, dtype=torch.float32) / 180. * np.pi
    sin, cos = torch.sin(ang), torch.cos(ang)
    cls = cls.sigmoid()
    rays = (rays * st.view(1, -1, 1)).clamp(min=FLOOR)
    valid = rays > 1
    x = rays * cos + pts[None, :, None, 0]
    y = rays * sin + pts[None, :, None, 1]
    box = torch.stack([x.min(-1)[0], y.min(-1)[0], x.max(-1)[0], y.max(-1)[0]], -1)
    return torch.cat((box, cls, x, y, valid.float()), -1).permute(0, 2, 1).contiguous()


def greedy_nms(boxes: np.ndarray, scores: np.ndarray, iou_thres: float):
    """Restatement of torchvision.ops.nms (torchvision 0.26.0, csrc/ops/cpu/nms_kernel.cpp): stable
    descending sort by score, greedy, suppress j when inter/(area_i+area_j-inter) > thr (strict).
    Returns (keep indices in score order, min |iou - thr| over the comparisons actually made)."""
    order = np.argsort(-scores.astype(np.float32), kind="stable")
    b = boxes.astype(np.float32)
    x1, y1, x2, y2 = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    area = (x2 - x1) * (y2 - y1)
    dead = np.zeros(len(order), dtype=bool)
    keep = []
    margin = np.inf
    thr = np.float32(iou_thres)
    for k, i in enumerate(order):
        if dead[k]:
            continue
        keep.append(i)
        rest = order[k + 1:]
        if rest.size == 0:
            break
        w = np.maximum(np.float32(0), np.minimum(x2[i], x2[rest]) - np.maximum(x1[i], x1[rest]))
        h = np.maximum(np.float32(0), np.minimum(y2[i], y2[rest]) - np.maximum(y1[i], y1[rest]))
        inter = w * h
        with np.errstate(invalid="ignore", divide="ignore"):
            iou = inter / (area[i] + area[rest] - inter)
        live = ~dead[k + 1:]
        if live.any():
            margin = min(margin, float(np.nanmin(np.abs(iou[live] - thr))))
        dead[k + 1:] |= iou > thr
    return np.array(keep, dtype=np.int64), margin


def nms(prediction, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False,
        multi_label=False, max_det=300, nc=0, max_nms=30000, max_wh=7680):
    """utils/ops.py:285-424 non_max_suppression (polar variant: boxes already xyxy).
    Returns (list of (n_i, 6+nm) tensors, min IoU-vs-threshold margin over the batch)."""
    assert 0 <= conf_thres <= 1 and 0 <= iou_thres <= 1
    if isinstance(prediction, (list, tuple)):
        prediction = prediction[0]
    bs = prediction.shape[0]
    nc = nc or (prediction.shape[1] - 4)
    nm = prediction.shape[1] - nc - 4
    mi = 4 + nc
    xc = prediction[:, 4:mi].amax(1) > conf_thres
    multi_label &= nc > 1
    prediction = prediction.transpose(-1, -2)
    output = [torch.zeros((0, 6 + nm))] * bs
    margin = math.inf
    for xi, x in enumerate(prediction):
        x = x[xc[xi]]
        if not x.shape[0]:
            continue
        box, cls, mask = x.split((4, nc, nm), 1)
        if multi_label:
            i, j = torch.where(cls > conf_thres)
            x = torch.cat((box[i], x[i, 4 + j, None], j[:, None].float(), mask[i]), 1)
        else:
            conf, j = cls.max(1, keepdim=True)
            x = torch.cat((box, conf, j.float(), mask), 1)[conf.view(-1) > conf_thres]
        if classes is not None:
            x = x[(x[:, 5:6] == torch.tensor(classes)).any(1)]
        n = x.shape[0]
        if not n:
            continue
        if n > max_nms:
            x = x[torch.sort(x[:, 4], descending=True, stable=True)[1][:max_nms]]
        c = x[:, 5:6] * (0 if agnostic else max_wh)
        boxes, scores = x[:, :4] + c, x[:, 4]
        keep, m = greedy_nms(boxes.numpy(), scores.numpy(), iou_thres)
        margin = min(margin, m)
        output[xi] = x[torch.from_numpy(keep[:max_det])]
    return output, margin


# --------------------------------------------------------------------------------------------
# dormant box terms: CIoU + DFL (BboxLoss)
# --------------------------------------------------------------------------------------------
def bbox_ciou(box1, box2, eps: float = 1e-7):
    """utils/metrics.py:77-130 bbox_iou(xywh=False, CIoU=True)"""
    b1_x1, b1_y1, b1_x2, b1_y2 = box1.chunk(4, -1)
    b2_x1, b2_y1, b2_x2, b2_y2 = box2.chunk(4, -1)
    w1, h1 = b1_x2 - b1_x1, b1_y2 - b1_y1 + eps
    w2, h2 = b2_x2 - b2_x1, b2_y2 - b2_y1 + eps
    inter = (torch.minimum(b1_x2, b2_x2) - torch.maximum(b1_x1, b2_x1)).clamp(0) * \
            (torch.minimum(b1_y2, b2_y2) - torch.maximum(b1_y1, b2_y1)).clamp(0)
    union = w1 * h1 + w2 * h2 - inter + eps
    iou = inter / union
    cw = torch.maximum(b1_x2, b2_x2) - torch.minimum(b1_x1, b2_x1)
    ch = torch.maximum(b1_y2, b2_y2) - torch.minimum(b1_y1, b2_y1)
    c2 = cw ** 2 + ch ** 2 + eps
    rho2 = ((b2_x1 + b2_x2 - b1_x1 - b1_x2) ** 2 + (b2_y1 + b2_y2 - b1_y1 - b1_y2) ** 2) / 4
    v = (4 / math.pi ** 2) * (torch.atan(w2 / h2) - torch.atan(w1 / h1)).pow(2)
    with torch.no_grad():
        alpha = v / (v - iou + (1 + eps))
    return iou - (rho2 / c2 + v * alpha)


def bbox_loss(pred_dist, pred_bboxes, anchor_points, target_bboxes, target_scores, target_scores_sum, fg_mask,
              reg_max: int = 15, use_dfl: bool = True):
    """utils/loss.py:61-87 BboxLoss.forward + _df_loss; bbox2dist utils/tal.py:1437-1440.
    reg_max here is BboxLoss.reg_max (= head reg_max - 1, i.e. 15)."""
    weight = target_scores.sum(-1)[fg_mask].unsqueeze(-1)
    iou = bbox_ciou(pred_bboxes[fg_mask], target_bboxes[fg_mask])
    loss_iou = ((1.0 - iou) * weight).sum() / target_scores_sum
    loss_dfl = torch.tensor(0.0)
    if use_dfl:
        x1y1, x2y2 = target_bboxes.chunk(2, -1)
        ltrb = torch.cat((anchor_points - x1y1, x2y2 - anchor_points), -1).clamp(0, reg_max - 0.01)
        pd = pred_dist[fg_mask].view(-1, reg_max + 1)
        tgt = ltrb[fg_mask]
        tl = tgt.long()
        tr = tl + 1
        wl = tr - tgt
        wr = 1 - wl
        ce = torch.nn.functional.cross_entropy
        dfl = (ce(pd, tl.view(-1), reduction="none").view(tl.shape) * wl +
               ce(pd, tr.view(-1), reduction="none").view(tl.shape) * wr).mean(-1, keepdim=True)
        loss_dfl = (dfl * weight).sum() / target_scores_sum
    return loss_iou, loss_dfl


# --------------------------------------------------------------------------------------------
# contour resampling (data-format step in front of the path)
# --------------------------------------------------------------------------------------------
def resample_segments(segments, n: int = 360):
    """utils/ops.py:676-693 resample_segments: close each (m,2) polygon, np.interp x and y onto
    linspace(0, m, n) (double precision), return float32 (n,2) arrays."""
    out = []
    for s in segments:
        s = np.asarray(s)
        s = np.concatenate((s, s[0:1, :]), axis=0)
        x = np.linspace(0, len(s) - 1, n)
        xp = np.arange(len(s))
        out.append(np.concatenate([np.interp(x, xp, s[:, i]) for i in range(2)], dtype=np.float32).reshape(2, -1).T)
    return out


# --------------------------------------------------------------------------------------------
# contour -> mask rasterisation
# --------------------------------------------------------------------------------------------
def _clip_line(w, h, p1, p2):
    """cv2.clipLine (OpenCV 4.x drawing.cpp): Cohen-Sutherland with double arithmetic truncated to integers."""
    x1, y1 = p1
    x2, y2 = p2
    right, bottom = w - 1, h - 1
    c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8
    c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8
    if (c1 & c2) == 0 and (c1 | c2) != 0:
        if c1 & 12:
            a = 0 if c1 < 8 else bottom
            x1 += int(float(a - y1) * (x2 - x1) / (y2 - y1))
            y1 = a
            c1 = (x1 < 0) + (x1 > right) * 2
        if c2 & 12:
            a = 0 if c2 < 8 else bottom
            x2 += int(float(a - y2) * (x2 - x1) / (y2 - y1))
            y2 = a
            c2 = (x2 < 0) + (x2 > right) * 2
        if (c1 & c2) == 0 and (c1 | c2) != 0:
            if c1:
                a = 0 if c1 == 1 else right
                y1 += int(float(a - x1) * (y2 - y1) / (x2 - x1))
                x1 = a
                c1 = 0
            if c2:
                a = 0 if c2 == 1 else right
                y2 += int(float(a - x2) * (y2 - y1) / (x2 - x1))
                x2 = a
                c2 = 0
    return (c1 | c2) == 0, (x1, y1), (x2, y2)


def _draw_line(img, p0, p1):
    """cv2.line(..., thickness 1, 8-connected): clipLine, then LineIterator left to right."""
    h, w = img.shape
    ok, p0, p1 = _clip_line(w, h, p0, p1)
    if not ok:
        return
    (x0, y0), (x1, y1) = p0, p1
    dx, dy = x1 - x0, y1 - y0
    if dx < 0:
        x0, y0, dx, dy = x1, y1, -dx, -dy
    sy = 1 if dy >= 0 else -1
    dy = abs(dy)
    steep = dy > dx
    a, b = (dy, dx) if steep else (dx, dy)
    err, x, y = a - 2 * b, x0, y0
    for _ in range(a + 1):
        if 0 <= x < w and 0 <= y < h:
            img[y, x] = 1
        m = err < 0
        err += -2 * b + (2 * a if m else 0)
        if steep:
            y += sy
            x += 1 if m else 0
        else:
            x += 1
            y += sy if m else 0


def fill_poly(points, h, w):
    """cv2.fillPoly(img, [points], color) for one integer polygon (the step utils/ops.py:794-809 has commented out):
    boundary lines plus even-odd scan-line fill in 16.16 fixed point (OpenCV drawing.cpp CollectPolyEdges /
    FillEdgeCollection): an edge covers scan lines y0 <= y < y1 with x = x0 + (y - y0) * dx, dx the truncated
    quotient; between the sorted crossings of a pair the pixels ceil(xa) .. floor(xb) are set.  -> (h,w) uint8 0/1.
    Pinned against cv2.fillPoly in tests/test_oracle_golden.py."""
    ONE = 1 << 16
    img = np.zeros((h, w), np.uint8)
    pts = [(int(p[0]), int(p[1])) for p in points]
    n = len(pts)
    if n == 0:
        return img
    edges = []
    for i in range(n):
        p0, p1 = pts[i - 1], pts[i]
        _draw_line(img, p0, p1)
        if p0[1] == p1[1]:
            continue
        if p0[1] < p1[1]:
            y0, y1, x = p0[1], p1[1], p0[0] * ONE
        else:
            y0, y1, x = p1[1], p0[1], p1[0] * ONE
        num, den = (p1[0] - p0[0]) * ONE, p1[1] - p0[1]
        q = abs(num) // abs(den)
        edges.append((y0, y1, x, q if (num >= 0) == (den > 0) else -q))
    if n < 3 and not edges:
        return img
    for y in range(0, h):
        xs = sorted(x + (y - y0) * d for (y0, y1, x, d) in edges if y0 <= y < y1)
        for k in range(0, len(xs) - 1, 2):
            x1, x2 = (xs[k] + ONE - 1) >> 16, xs[k + 1] >> 16
            if x1 < w and x2 >= 0:
                x1, x2 = max(x1, 0), min(x2, w - 1)
                if x1 <= x2:
                    img[y, x1:x2 + 1] = 1
    return img


def contour_masks(rows, R, h, w):
    """The intended ops.process_mask of the polar fork (utils/ops.py:768-825 with the commented loop restored):
    rows (n, 6+3R) NMS output; per detection the valid contour points, truncated to int32, filled -> (n,h,w) uint8."""
    rows = np.asarray(rows, np.float32)
    out = np.zeros((rows.shape[0], h, w), np.uint8)
    for i, r in enumerate(rows):
        xx, yy, ok = r[6:6 + R], r[6 + R:6 + 2 * R], r[6 + 2 * R:6 + 3 * R] != 0
        pts = list(zip(xx[ok].astype(np.int32).tolist(), yy[ok].astype(np.int32).tolist()))
        out[i] = fill_poly(pts, h, w)
    return out


def mask_iou(m1, m2, eps=1e-7):
    """utils/metrics.py:133-155 mask_iou: (N,n) x (M,n) 0/1 masks -> (N,M) IoU in fp32."""
    a = torch.as_tensor(m1).float()
    b = torch.as_tensor(m2).float()
    inter = torch.matmul(a, b.t()).clamp(0)
    union = (a.sum(1)[:, None] + b.sum(1)[None]) - inter
    return inter / (union + eps)
