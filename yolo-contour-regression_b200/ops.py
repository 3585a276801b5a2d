"""Host-side mirror of `ultralytics/utils/ops.py::non_max_suppression` (polar variant,
utils/ops.py:285-424; paths relative to /root/reference/ultralytics-main/ultralytics/)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L


def non_max_suppression(
        prediction,
        conf_thres=0.25,
        iou_thres=0.45,
        classes=None,
        agnostic=False,
        multi_label=False,
        labels=(),
        max_det=300,
        nc=0,
        max_time_img=0.05,
        max_nms=30000,
        max_wh=7680,
):
    """Same 12-argument signature, assertions and return type as utils/ops.py:285-298:
    list (one per image) of (n_i, 6+nm) tensors [x1,y1,x2,y2,conf,cls,masks...] in descending score order.
    Boxes are already xyxy (the polar variant does not convert).  One filter kernel, one per-image sort
    kernel and one per-image suppression+gather kernel replace the per-image Python loop; the only host
    synchronisation is the read of the B kept-counts that the list-of-tensors return type requires.
    `max_time_img` is accepted and ignored (no wall-clock bail-out).  `max_det` is limited to 1024 (ValueError
    beyond).  Apriori `labels` (save_hybrid autolabelling, utils/ops.py:368-374): the label rows of image i
    ([class, x1, y1, x2, y2] each) join that image's candidates with score 1.0 for their class and zero masks,
    as the reference intends (its own `v` is one column too wide for the polar layout, utils/ops.py:370, and
    raises in torch.cat)."""
    assert 0 <= conf_thres <= 1, f'Invalid Confidence threshold {conf_thres}, valid values are between 0.0 and 1.0'
    assert 0 <= iou_thres <= 1, f'Invalid IoU {iou_thres}, valid values are between 0.0 and 1.0'
    if isinstance(prediction, (list, tuple)):
        prediction = prediction[0]
    L.require_cuda(prediction)
    dev = prediction.device
    pred = prediction if (prediction.dtype == torch.float32 and prediction.is_contiguous()) \
        else prediction.float().contiguous()
    B, CH, A = pred.shape
    nc = nc or (CH - 4)
    nm = CH - nc - 4
    if max_det > 1024:
        raise ValueError(f"max_det={max_det}: at most 1024 detections per image are supported")
    if B == 0 or (A == 0 and not (labels and any(len(lb) for lb in labels))):
        return [torch.zeros((0, 6 + nm), device=dev)] * B      # utils/ops.py:362: nothing to look at
    if labels and any(len(lb) for lb in labels):
        # apriori labels become extra candidate columns: box, a one-hot class score of 1.0, zero masks
        n_extra = max(len(lb) for lb in labels)
        extra = torch.zeros(B, CH, n_extra, device=dev, dtype=torch.float32)
        for xi, lb in enumerate(labels):
            if len(lb):
                lb = torch.as_tensor(lb, dtype=torch.float32, device=dev).reshape(-1, 5)
                k = torch.arange(lb.shape[0], device=dev)
                extra[xi, :4, :lb.shape[0]] = lb[:, 1:5].t()
                extra[xi, 4 + lb[:, 0].long(), k] = 1.0
        pred = torch.cat((pred, extra), 2).contiguous()
        A = pred.shape[2]
    cfg = L.NmsCfg()
    cfg.conf_thres, cfg.iou_thres = float(conf_thres), float(iou_thres)
    cfg.agnostic, cfg.multi_label = int(bool(agnostic)), int(bool(multi_label) and nc > 1)
    cfg.max_det, cfg.nc, cfg.max_nms, cfg.max_wh = int(max_det), int(nc), int(max_nms), float(max_wh)
    cls_t = None
    if classes is not None:
        cls_t = torch.as_tensor(list(classes), dtype=torch.int32, device=dev)
        cfg.classes, cfg.n_classes = cls_t.data_ptr(), cls_t.numel()
    else:
        cfg.classes, cfg.n_classes = None, 0
    lib = L.lib()
    nbytes = lib.ycr_nms_workspace_bytes(B, A, CH, C.byref(cfg))
    ws = L.Workspace.get("nms", nbytes, dev)
    cfg.compact_rows = 1   # rows of all images back to back: one split instead of B Python slices
    hint = getattr(prediction, "_ycr_best_class", None)   # left by head.decode on its own output tensor
    best_t = None
    if (hint is not None and pred is prediction and not cfg.multi_label and hint[1] == prediction._version
            and hint[2] == nc and tuple(hint[0].shape) == (B, A, 2)):
        best_t = hint[0]
        cfg.best_class = best_t.data_ptr()
    else:
        cfg.best_class = None
    fhint = getattr(prediction, "_ycr_feats", None)
    keep_feats = None
    if fhint is not None and pred is prediction and hint is not None and hint[1] == prediction._version and nm == 3 * fhint[3]:
        keep_feats, cgrid = fhint[0], fhint[1]     # (same tensor, untouched since decode wrote it)
        for li, f in enumerate(keep_feats):
            cfg.feats[li] = f.data_ptr()
        cfg.grid = C.pointer(cgrid)
        cfg.feats_dtype, cfg.rays = fhint[2], fhint[3]
    rows = torch.empty(B * max_det, 6 + nm, device=dev, dtype=torch.float32)
    counts = torch.empty(B, device=dev, dtype=torch.int32)
    rc = lib.ycr_nms(pred.data_ptr(), B, CH, A, C.byref(cfg), rows.data_ptr(), counts.data_ptr(), ws.data_ptr(),
                     ws.numel(), L.stream_ptr(dev))
    L.check(rc, "ycr_nms")
    # the list-of-tensors return type needs the kept counts on the host: one asynchronous copy into pinned memory
    # and one event wait (the only synchronisation of the call)
    n = _counts_to_host(counts)
    del cls_t, best_t, keep_feats
    return list(torch.split(rows[:sum(n)], n))


def detect(feats, strides, nc, rays=36, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False, max_det=300,
           max_nms=30000, max_wh=7680, packed=False):
    """Deployment form of `Segment.forward(eval)` + `non_max_suppression` (single-label, the predictor's call): the
    head feature maps (fp32 / fp16 / bf16, the export branch's raw outputs nn/modules/head.py:572-574) go to the
    kept rows in one library call that never writes the (B, 4+nc+3R, A) prediction tensor.  Same list of
    (n_i, 6+3R) tensors, bit for bit, as the two-call form.  `packed=True` returns what the library call itself
    produces instead - one `(sum n_i, 6+3R)` tensor with the images' rows back to back and the device tensor of the
    B kept counts - without reading the counts on the host (no synchronisation, no B row views)."""
    assert 0 <= conf_thres <= 1, f'Invalid Confidence threshold {conf_thres}, valid values are between 0.0 and 1.0'
    assert 0 <= iou_thres <= 1, f'Invalid IoU {iou_thres}, valid values are between 0.0 and 1.0'
    L.require_cuda(*feats)
    dev = feats[0].device
    dt = feats[0].dtype if feats[0].dtype in L.DTYPE_CODE else torch.float32
    feats = [f if (f.dtype == dt and f.is_contiguous()) else f.to(dt).contiguous() for f in feats]
    B = feats[0].shape[0]
    if feats[0].shape[1] != rays + nc:
        raise ValueError(f"feature maps have {feats[0].shape[1]} channels, expected {rays + nc}")
    if max_det > 1024:
        raise ValueError(f"max_det={max_det}: at most 1024 detections per image are supported")
    if B == 0:
        return (torch.zeros(0, 6 + 3 * rays, device=dev), torch.zeros(0, dtype=torch.int32, device=dev)) if packed else []
    cgrid = L.make_grid([tuple(f.shape[2:]) for f in feats], [float(s) for s in strides])
    cfg = L.NmsCfg()
    cfg.conf_thres, cfg.iou_thres = float(conf_thres), float(iou_thres)
    cfg.agnostic, cfg.multi_label = int(bool(agnostic)), 0
    cfg.max_det, cfg.nc, cfg.max_nms, cfg.max_wh = int(max_det), int(nc), int(max_nms), float(max_wh)
    cls_t = None
    if classes is not None:
        cls_t = torch.as_tensor(list(classes), dtype=torch.int32, device=dev)
        cfg.classes, cfg.n_classes = cls_t.data_ptr(), cls_t.numel()
    cfg.compact_rows = 1
    lib = L.lib()
    nbytes = lib.ycr_detect_workspace_bytes(C.byref(cgrid), B, C.byref(cfg))
    ws = L.Workspace.get("detect", nbytes, dev)
    rows = torch.empty(B * max_det, 6 + 3 * rays, device=dev, dtype=torch.float32)
    counts = torch.empty(B, device=dev, dtype=torch.int32)
    rc = lib.ycr_detect(C.byref(cgrid), L.ptr_array(feats), L.DTYPE_CODE[dt], B, int(nc), int(rays), C.byref(cfg),
                        rows.data_ptr(), counts.data_ptr(), ws.data_ptr(), ws.numel(), L.stream_ptr(dev))
    L.check(rc, "ycr_detect")
    if packed:
        return rows, counts   # rows past counts.sum() are unspecified
    n = _counts_to_host(counts)
    del cls_t
    return list(torch.split(rows[:sum(n)], n))


_PINNED_COUNTS: dict = {}


def _counts_to_host(counts):
    B = counts.numel()
    key = (B, str(counts.device))
    buf = _PINNED_COUNTS.get(key)
    if buf is None:
        buf = _PINNED_COUNTS[key] = torch.empty(B, dtype=torch.int32).pin_memory()
    buf.copy_(counts, non_blocking=True)
    ev = torch.cuda.Event()
    ev.record(torch.cuda.current_stream(counts.device))
    ev.synchronize()
    return buf.tolist()


def resample_segments(segments, n=1000, device=None):
    """utils/ops.py:676-693 — list of (m_i, 2) polygons (numpy arrays or tensors) -> list of (n, 2) float32
    tensors on the device, each polygon closed and linearly resampled to n points; values are bit-identical
    to the reference's numpy implementation.  One kernel launch for the whole list."""
    import numpy as np
    if len(segments) == 0:
        return []
    ts = [torch.as_tensor(np.asarray(s) if not torch.is_tensor(s) else s, dtype=torch.float32).reshape(-1, 2)
          for s in segments]
    dev = torch.device(device) if device is not None else next((t.device for t in ts if t.is_cuda), torch.device("cuda"))
    if dev.type != "cuda":
        raise L.YcrError("ycr_b200 kernels need a CUDA device; there is no CPU path")
    offs = torch.zeros(len(ts) + 1, dtype=torch.int32)
    offs[1:] = torch.cumsum(torch.tensor([t.shape[0] for t in ts]), 0)
    pts = torch.cat([t.to(dev) for t in ts]).contiguous()
    offs_d = offs.to(dev)
    out = torch.empty(len(ts), int(n), 2, device=dev, dtype=torch.float32)
    rc = L.lib().ycr_resample_segments(pts.data_ptr(), offs_d.data_ptr(), len(ts), int(n), out.data_ptr(), L.stream_ptr(dev))
    L.check(rc, "ycr_resample_segments")
    return list(out.unbind(0))


def process_mask(protos, masks_in, bboxes, shape, upsample=False):
    """utils/ops.py:768-825 (polar variant) with the fill loop it has commented out (:794-809) done on the device:
    `masks_in` (n, 3R) are the mask columns of NMS rows [x_0.. | y_0.. | valid_0..]; per detection the valid contour
    points, truncated to int32, are filled as cv2.fillPoly does -> (n, h, w) uint8 masks with values 0/1
    (float32 when `upsample`, the predictor's call, which feeds them to Results).  `protos` and `bboxes` are
    accepted and unused, as in the reference."""
    L.require_cuda(masks_in)
    n, pn = masks_in.shape
    R = pn // 3
    h, w = int(shape[0]), int(shape[1])
    dev = masks_in.device
    rows = torch.empty(n, 6 + 3 * R, device=dev, dtype=torch.float32)
    rows[:, 6:] = masks_in.float()
    masks = torch.empty(n, h, w, device=dev, dtype=torch.uint8)
    rc = L.lib().ycr_rasterize_contours(rows.data_ptr(), rows.stride(0), n, R, h, w, masks.data_ptr(), L.stream_ptr(dev))
    L.check(rc, "ycr_rasterize_contours")
    return masks.float() if upsample else masks


def rasterize_rows(rows, R, shape):
    """Same, straight from NMS rows (n, 6+3R) - no copy of the mask columns."""
    L.require_cuda(rows)
    rows = rows if (rows.dtype == torch.float32 and rows.stride(1) == 1) else rows.float().contiguous()
    n = rows.shape[0]
    h, w = int(shape[0]), int(shape[1])
    masks = torch.empty(n, h, w, device=rows.device, dtype=torch.uint8)
    rc = L.lib().ycr_rasterize_contours(rows.data_ptr(), rows.stride(0), n, int(R), h, w, masks.data_ptr(),
                                        L.stream_ptr(rows.device))
    L.check(rc, "ycr_rasterize_contours")
    return masks


def mask_iou(mask1, mask2, eps=1e-7):
    """utils/metrics.py:133-155: mask1 (N, n) ground-truth masks, mask2 (M, n) predicted masks (uint8 or float,
    non-zero = set) -> (N, M) IoU.  Both sets are bit-packed and intersected with popcounts instead of the
    reference's (N, n) x (n, M) float matmul."""
    L.require_cuda(mask1, mask2)
    dev = mask1.device

    def prep(m):
        if m.dtype == torch.bool:
            m = m.to(torch.uint8)
        if m.dtype not in (torch.uint8, torch.float32):
            m = m.float()
        return m.contiguous(), (0 if m.dtype == torch.uint8 else 1)
    a, da = prep(mask1)
    b, db = prep(mask2)
    N, n = a.shape
    M = b.shape[0]
    out = torch.zeros(N, M, device=dev, dtype=torch.float32)
    if N == 0 or M == 0:
        return out
    lib = L.lib()
    nbytes = lib.ycr_mask_iou_workspace_bytes(N, M, n)
    ws = L.Workspace.get("mask_iou", nbytes, dev)
    rc = lib.ycr_mask_iou(a.data_ptr(), da, b.data_ptr(), db, N, M, n, float(eps), out.data_ptr(), ws.data_ptr(),
                          ws.numel(), L.stream_ptr(dev))
    L.check(rc, "ycr_mask_iou")
    return out
