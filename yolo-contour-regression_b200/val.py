"""Boundary mirror of the segment validator hooks (models/yolo/segment/val.py:46-61 `postprocess`,
:149-219 `update_metrics`, :226-261 `_process_batch`; paths relative to
/root/reference/ultralytics-main/ultralytics/).  SURVEY.md §2 marks these "boundary only": signatures are
kept, the kernel behind them is the batched NMS.  Metric bookkeeping is the reference's CPU/numpy matching,
restated; predicted masks are all-zero in the reference snapshot (ops.process_mask polar variant,
utils/ops.py:768-825), so mask correctness is all False here as well."""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np
import torch

from .ops import non_max_suppression


def box_iou(box1, box2, eps=1e-7):
    """utils/metrics.py:56-74 — (N,4) x (M,4) xyxy -> (N,M)."""
    (a1, a2), (b1, b2) = box1.unsqueeze(1).chunk(2, 2), box2.unsqueeze(0).chunk(2, 2)
    inter = (torch.min(a2, b2) - torch.max(a1, b1)).clamp_(0).prod(2)
    return inter / ((a2 - a1).prod(2) + (b2 - b1).prod(2) - inter + eps)


class SegmentationValidator:
    """`postprocess(preds)` and `update_metrics(preds, batch)` with the reference's semantics.

    args: conf (0.001 in val, engine/validator.py:85), iou 0.7, max_det 300, single_cls, overlap_mask."""

    def __init__(self, nc=80, args=None, device="cuda"):
        self.nc = nc
        self.args = args or SimpleNamespace(conf=0.001, iou=0.7, max_det=300, single_cls=False, overlap_mask=True)
        self.device = torch.device(device)
        self.lb = []
        self.iouv = torch.linspace(0.5, 0.95, 10, device=self.device)
        self.niou = self.iouv.numel()
        self.stats = []
        self.seen = 0

    def postprocess(self, preds):
        """models/yolo/segment/val.py:46-61: NMS with multi_label=True (without the stray print)."""
        return non_max_suppression(preds[0] if isinstance(preds, (list, tuple)) else preds, self.args.conf,
                                   self.args.iou, labels=self.lb, multi_label=True,
                                   agnostic=self.args.single_cls, max_det=self.args.max_det, nc=self.nc)

    def _process_batch(self, detections, labels):
        """models/yolo/segment/val.py:226-261, box branch."""
        iou = box_iou(labels[:, 1:], detections[:, :4])
        correct = np.zeros((detections.shape[0], self.iouv.shape[0])).astype(bool)
        correct_class = labels[:, 0:1] == detections[:, 5]
        for i in range(len(self.iouv)):
            x = torch.where((iou >= self.iouv[i]) & correct_class)
            if x[0].shape[0]:
                matches = torch.cat((torch.stack(x, 1), iou[x[0], x[1]][:, None]), 1).cpu().numpy()
                if x[0].shape[0] > 1:
                    matches = matches[matches[:, 2].argsort()[::-1]]
                    matches = matches[np.unique(matches[:, 1], return_index=True)[1]]
                    matches = matches[np.unique(matches[:, 0], return_index=True)[1]]
                correct[matches[:, 1].astype(int), i] = True
        return torch.tensor(correct, dtype=torch.bool, device=detections.device)

    def update_metrics(self, preds, batch):
        """models/yolo/segment/val.py:149-219 for letterbox-free batches (ratio_pad identity)."""
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
            predn = pred.clone()
            if nl:
                xy, wh = bbox[:, :2], bbox[:, 2:] / 2
                tbox = torch.cat((xy - wh, xy + wh), 1) * torch.tensor((width, height, width, height),
                                                                      device=self.device)
                labelsn = torch.cat((cls, tbox), 1)
                correct_bboxes = self._process_batch(predn, labelsn)
            self.stats.append((correct_bboxes, correct_masks, pred[:, 4], pred[:, 5], cls.squeeze(-1)))
