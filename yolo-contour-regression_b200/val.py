"""Segment validator hooks of the polar path (models/yolo/segment/val.py:46-61 `postprocess`, :149-219
`update_metrics`, :226-261 `_process_batch`; paths relative to /root/reference/ultralytics-main/ultralytics/).

With `install()` the reference's own SegmentationValidator already runs on the rebound kernels (NMS, contour
rasterisation, mask IoU).  This class is the same pair of hooks for callers that do not have an `ultralytics`
checkout: everything stays on the device - batched NMS, `process_mask` = contour fill, `mask_iou` = bit-packed
popcounts, and the detection-to-label matching as a handful of tensor ops instead of the reference's numpy loop."""
from __future__ import annotations

from types import SimpleNamespace

import torch
import torch.nn.functional as F

from . import ops


def match_predictions(iou, correct_class, iouv):
    """The matching rule of `_process_batch` (models/yolo/segment/val.py:247-261) without the host round trip.
    iou (L, D) labels x detections, correct_class (L, D) bool, iouv (T,) thresholds -> correct (D, T) bool.

    Per threshold, among the pairs with iou >= threshold and equal class: every detection keeps its best label
    (highest IoU), then every label keeps the first (lowest-index) detection that chose it - what the reference's
    sort / np.unique(det) / np.unique(label) sequence leaves."""
    L_, D = iou.shape
    T = iouv.numel()
    correct = torch.zeros(D, T, dtype=torch.bool, device=iou.device)
    if L_ == 0 or D == 0:
        return correct
    det = torch.arange(D, device=iou.device)
    for i in range(T):
        ok = (iou >= iouv[i]) & correct_class
        has = ok.any(0)                                           # detections with at least one candidate label
        best = torch.where(ok, iou, iou.new_full((), -1.0)).argmax(0)   # best label per detection
        # first detection per label among those that chose it
        first = torch.full((L_,), D, device=iou.device, dtype=torch.long)
        first.scatter_reduce_(0, best[has], det[has], reduce="amin", include_self=True)
        keep = first[first < D]
        correct[keep, i] = True
    return correct


def _box_iou(a, b, eps=1e-7):
    """(L,4) x (D,4) xyxy boxes -> (L,D) IoU (what utils/metrics.py box_iou computes)."""
    lt = torch.maximum(a[:, None, :2], b[None, :, :2])
    rb = torch.minimum(a[:, None, 2:], b[None, :, 2:])
    inter = (rb - lt).clamp(min=0).prod(2)
    area_a = (a[:, 2:] - a[:, :2]).prod(1)
    area_b = (b[:, 2:] - b[:, :2]).prod(1)
    return inter / (area_a[:, None] + area_b[None] - inter + eps)


class SegmentationValidator:
    """`postprocess(preds)` and `update_metrics(preds, batch)` with the reference's semantics (letterbox-free
    batches: ratio_pad identity).

    args: conf (0.001 in val, engine/validator.py:85), iou 0.7, max_det 300, single_cls, overlap_mask."""

    def __init__(self, nc=80, args=None, device="cuda", rays=36):
        self.nc = nc
        self.rays = rays
        self.args = args or SimpleNamespace(conf=0.001, iou=0.7, max_det=300, single_cls=False, overlap_mask=True)
        self.device = torch.device(device)
        self.lb = []
        self.iouv = torch.linspace(0.5, 0.95, 10, device=self.device)
        self.niou = self.iouv.numel()
        self.stats = []
        self.seen = 0

    def postprocess(self, preds):
        """models/yolo/segment/val.py:46-61: NMS with multi_label=True."""
        return ops.non_max_suppression(preds[0] if isinstance(preds, (list, tuple)) else preds, self.args.conf,
                                       self.args.iou, labels=self.lb, multi_label=True,
                                       agnostic=self.args.single_cls, max_det=self.args.max_det, nc=self.nc)

    def _gt_masks(self, batch, si, idx, nl, shape):
        """Per-label 0/1 masks at the prediction resolution (models/yolo/segment/val.py:236-243)."""
        m = batch["masks"].to(self.device).float()
        if getattr(self.args, "overlap_mask", True):
            index = torch.arange(nl, device=self.device).view(nl, 1, 1) + 1
            gt = (m[si][None] == index).float()
        else:
            gt = m[idx]
        if gt.shape[1:] != tuple(shape):
            gt = F.interpolate(gt[None], shape, mode="bilinear", align_corners=False)[0].gt_(0.5)
        return gt

    def update_metrics(self, preds, batch):
        """models/yolo/segment/val.py:149-219."""
        height, width = batch["img"].shape[2:]
        for si, pred in enumerate(preds):
            idx = batch["batch_idx"] == si
            cls = batch["cls"][idx].to(self.device)
            bbox = batch["bboxes"][idx].to(self.device)
            nl, npr = cls.shape[0], pred.shape[0]
            correct_masks = torch.zeros(npr, self.niou, dtype=torch.bool, device=self.device)
            correct_bboxes = torch.zeros(npr, self.niou, dtype=torch.bool, device=self.device)
            self.seen += 1
            if npr == 0:
                if nl:
                    self.stats.append((correct_bboxes, correct_masks, *torch.zeros((2, 0), device=self.device),
                                       cls.squeeze(-1)))
                continue
            if self.args.single_cls:
                pred[:, 5] = 0
            if nl:
                xy, wh = bbox[:, :2], bbox[:, 2:] / 2
                tbox = torch.cat((xy - wh, xy + wh), 1) * torch.tensor((width, height, width, height),
                                                                      device=self.device)
                same = cls.view(-1, 1) == pred[:, 5].view(1, -1)
                correct_bboxes = match_predictions(_box_iou(tbox, pred[:, :4]), same, self.iouv)
                if "masks" in batch:
                    pm = ops.rasterize_rows(pred, self.rays, (height, width))
                    gm = self._gt_masks(batch, si, idx, nl, (height, width))
                    iou = ops.mask_iou(gm.view(nl, -1), pm.view(npr, -1))
                    correct_masks = match_predictions(iou, same, self.iouv)
            self.stats.append((correct_bboxes, correct_masks, pred[:, 4], pred[:, 5], cls.squeeze(-1)))
