"""Drop the B200 kernels into a stock checkout of the reference: `ycr_b200.install.install()` rebinds
the hot-path symbols of an importable `ultralytics` package so that
YOLO('yolov8n-seg.yaml').train()/val()/predict() run unchanged on top of them.  See INTEGRATION.md."""
from __future__ import annotations

import importlib

PATCHES = (
    # (module, attribute, replacement module in this package, attribute)
    ("ultralytics.utils.tal", "TaskAlignedAssigner", "tal", "TaskAlignedAssigner"),
    ("ultralytics.utils.tal", "make_anchors_polar", "tal", "make_anchors_polar"),
    ("ultralytics.utils.tal", "MaskIOU", "tal", "MaskIOU"),
    ("ultralytics.utils.loss", "v8SegmentationLoss", "loss", "v8SegmentationLoss"),
    ("ultralytics.utils.loss", "MaskIOULoss", "loss", "MaskIOULoss"),
    ("ultralytics.utils.loss", "TaskAlignedAssigner", "tal", "TaskAlignedAssigner"),
    ("ultralytics.nn.tasks", "v8SegmentationLoss", "loss", "v8SegmentationLoss"),
    ("ultralytics.utils.ops", "non_max_suppression", "ops", "non_max_suppression"),
    ("ultralytics.utils.ops", "process_mask", "ops", "process_mask"),
    ("ultralytics.utils.metrics", "mask_iou", "ops", "mask_iou"),
    ("ultralytics.models.yolo.segment.val", "mask_iou", "ops", "mask_iou"),
    ("ultralytics.nn.modules.head", "Segment", "head", "Segment"),
    ("ultralytics.nn.modules", "Segment", "head", "Segment"),
    ("ultralytics.nn.tasks", "Segment", "head", "Segment"),
)


def install(strict: bool = False):
    """Returns the list of (module, attribute) pairs that were rebound.  Modules that cannot be imported
    are skipped unless strict=True."""
    done = []
    for mod_name, attr, ours_mod, ours_attr in PATCHES:
        try:
            mod = importlib.import_module(mod_name)
        except Exception:
            if strict:
                raise
            continue
        ours = importlib.import_module(f"ycr_b200.{ours_mod}")
        if hasattr(mod, attr):
            setattr(mod, attr, getattr(ours, ours_attr))
            done.append((mod_name, attr))
    return done
