"""Shared by the stock-flow test: drive the reference's own model / validator / predictor objects (baseline/_ref)
on the BASELINE configs[0] batch (2 x 3 x 640 x 640, 8 GTs per image, 36 rays)."""
import os
import shutil
import tempfile
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch

from util import synth


def c1_batch(nc=10, seed=1):
    cfg = synth.PathConfig("C1", 2, 8, 640, nc=nc)
    batch = synth.make_gts(cfg, seed)
    g = torch.Generator().manual_seed(seed)
    batch["img"] = torch.rand(2, 3, 640, 640, generator=g)
    # overlap-format GT masks (index map, 1..n per image) from the same contours, at mask_ratio 4
    import cv2
    masks = np.zeros((2, 160, 160), np.uint8)
    for b in range(2):
        for k, seg in enumerate(batch["segments"][b]):
            cv2.fillPoly(masks[b], [(seg.numpy() * 160).astype(np.int32).reshape(-1, 1, 2)], k + 1)
    batch["masks"] = torch.from_numpy(masks)
    batch["ori_shape"] = [(640, 640), (640, 640)]
    batch["ratio_pad"] = [((1.0, 1.0), (0.0, 0.0))] * 2
    batch["im_file"] = ["a.jpg", "b.jpg"]
    return cfg, batch


def build_model(nc=10):
    """SegmentationModel from the reference's own yolov8-seg.yaml (scale n), with the hyper-parameters the loss reads."""
    from baseline import refload
    from ultralytics.nn.tasks import SegmentationModel
    d = tempfile.mkdtemp(prefix="ycr_flow_")
    y = os.path.join(d, "yolov8n-seg.yaml")
    shutil.copy(refload.model_yaml(), y)
    m = SegmentationModel(y, ch=3, nc=nc, verbose=False)
    m.args = SimpleNamespace(box=7.5, cls=0.5, dfl=1.5, overlap_mask=True)
    return m


def make_validator(model, device):
    from ultralytics.cfg import get_cfg
    from ultralytics.utils import DEFAULT_CFG
    from ultralytics.models.yolo.segment.val import SegmentationValidator
    args = get_cfg(DEFAULT_CFG, dict(conf=0.001, iou=0.7, max_det=300, plots=False, save_json=False, task="segment"))
    v = SegmentationValidator(save_dir=Path(tempfile.mkdtemp(prefix="ycr_val_")), args=args)
    v.device = torch.device(device)
    v.data = {"val": ""}
    v.training = False
    v.init_metrics(model)
    v.iouv = torch.linspace(0.5, 0.95, 10, device=v.device)
    v.niou = v.iouv.numel()
    v.lb = []
    v.batch_i = 0
    return v


def make_predictor(model, device):
    from ultralytics.models.yolo.segment.predict import SegmentationPredictor
    p = SegmentationPredictor(overrides=dict(conf=0.25, iou=0.7, max_det=300, save=False, verbose=False))
    p.model = SimpleNamespace(names=model.names)
    p.batch = (["a.jpg", "b.jpg"], None, None, None)
    return p
