"""Host-side mirror of `ultralytics/utils/tal.py` for the polar path (reference file:line cited per
symbol; paths relative to /root/reference/ultralytics-main/ultralytics/).  Same names, same argument
meaning, same error behaviour; all device work goes through the C-ABI library."""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _lib as L


def make_anchors_polar(feats, strides, grid_cell_offset=0.5):
    """utils/tal.py:1393-1407 — anchor centres (grid units), stride column, per-level stride list.
    Plain tensor plumbing (a few KB); the kernels regenerate anchors analytically."""
    anchor_points, stride_tensor = [], []
    dtype, device = feats[0].dtype, feats[0].device
    for i, stride in enumerate(strides):
        _, _, h, w = feats[i].shape
        sx = torch.arange(end=w, device=device, dtype=dtype) + grid_cell_offset
        sy = torch.arange(end=h, device=device, dtype=dtype) + grid_cell_offset
        sy, sx = torch.meshgrid(sy, sx, indexing="ij")
        anchor_points.append(torch.stack((sx, sy), -1).view(-1, 2))
        stride_tensor.append(torch.full((h * w, 1), float(stride), dtype=dtype, device=device))
    return torch.cat(anchor_points), torch.cat(stride_tensor), stride_tensor


def _rows(t: torch.Tensor, width: int):
    """(B,G,width) possibly a split view of the packed (B,G,725) tensor -> (tensor, row_stride)."""
    if t.dim() != 3 or t.shape[2] != width:
        raise ValueError(f"expected (B,G,{width}) got {tuple(t.shape)}")
    if t.dtype != torch.float32:
        t = t.float()
    G = t.shape[1]
    ok = t.stride(2) == 1 and (t.shape[0] == 1 or t.stride(0) == G * t.stride(1)) if G > 0 else True
    if not ok:
        t = t.contiguous()
    return t, (t.stride(1) if G > 0 else width)


def gt_struct(gt_labels, gt_bboxes, gt_coor, mask_gt=None):
    keep = []
    g = L.Gt()
    g.B, g.G = gt_bboxes.shape[0], gt_bboxes.shape[1]
    for name, t, w in (("labels", gt_labels, 1), ("boxes", gt_bboxes, 4), ("coor", gt_coor, 2 * L.CONTOUR_POINTS)):
        t, rs = _rows(t, w)
        keep.append(t)
        setattr(g, name, t.data_ptr())
        setattr(g, name + "_stride", rs)
    if mask_gt is not None:
        t, rs = _rows(mask_gt, 1)
        keep.append(t)
        g.mask_gt, g.mask_stride = t.data_ptr(), rs
    else:
        g.mask_gt, g.mask_stride = None, 0
    return g, keep


class TaskAlignedAssigner(nn.Module):
    """utils/tal.py:1109-1390 (polar TaskAlignedAssigner).  `forward` keeps the 11-argument signature of
    utils/tal.py:1135 and returns the 8-tuple of utils/tal.py:1204.

    Deviations (documented): an all-empty batch (G == 0) returns a consistent 8-tuple instead of the
    reference's 6-tuple that its only caller cannot unpack; ties are broken lowest-index; anchors must
    be the make_anchors_polar grid (the kernels enumerate in-box anchors analytically)."""

    def __init__(self, topk=13, num_classes=80, alpha=1.0, beta=3.0, eps=1e-9):
        super().__init__()
        self.topk = topk
        self.num_classes = num_classes
        self.bg_idx = num_classes
        self.alpha = alpha
        self.beta = beta
        self.eps = eps
        self.debug_metrics = False  # also return dense overlaps / align_metric (tests)
        self._grid_cache = {}

    def _grid_from_anchors(self, anc_points, stride_tensor, A):
        """Level shapes recovered from the anchors themselves (when the caller passes neither `ss`/`imgsz` nor
        `grid`): a level is a run of equal strides, its width the length of the first run of increasing x."""
        st = stride_tensor.detach().reshape(-1).float().cpu()
        xs = anc_points.detach()[:, 0].float().cpu()
        if st.numel() != A or xs.numel() != A:
            raise ValueError("anc_points / stride_tensor do not describe the anchors of pd_scores")
        cuts = [0] + (torch.nonzero(st[1:] != st[:-1]).flatten() + 1).tolist() + [A]
        # (no cache: two grids with the same anchor counts per level can differ in width, e.g. 80x60 and 60x80)
        shapes, strides = [], []
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            x = xs[lo:hi]
            wrap = torch.nonzero(x[1:] <= x[:-1]).flatten()
            w = int(wrap[0]) + 1 if wrap.numel() else hi - lo
            if (hi - lo) % w:
                raise ValueError("anchor layout is not the make_anchors_polar grid")
            shapes.append(((hi - lo) // w, w))
            strides.append(float(st[lo]))
        return shapes, strides

    def _grid(self, ss, imgsz, A):
        hw = [float(v) for v in (imgsz.tolist() if torch.is_tensor(imgsz) else imgsz)]
        key = (A, tuple(int(s.shape[0]) for s in ss), tuple(hw))   # the image size fixes the level widths
        g = self._grid_cache.get(key)
        if g is None:
            strides = [float(s.flatten()[0]) for s in ss]
            shapes = []
            for s, st in zip(ss, strides):
                w = int(round(hw[1] / st))
                n = int(s.shape[0])
                if w <= 0 or n % w:
                    raise ValueError("anchor layout is not the make_anchors_polar grid")
                shapes.append((n // w, w))
            if sum(h * w for h, w in shapes) != A:
                raise ValueError("anchor count does not match the level shapes")
            g = (shapes, strides)
            self._grid_cache[key] = g
        return g

    @torch.no_grad()
    def forward(self, pd_scores, pd_bboxes, anc_points, gt_labels, gt_bboxes, mask_gt, gt_coor, stride_tensor, ss,
                gt_center, imgsz, grid=None):
        """Arguments as utils/tal.py:1135 (`anc_points`, `stride_tensor`, `gt_center` do not affect the
        result there either).  `grid=(level_shapes, strides)` is an optional extension that skips the
        host round-trip used to recover the level shapes from `ss` and `imgsz`."""
        L.require_cuda(pd_scores, pd_bboxes, gt_bboxes, gt_coor)
        dev = pd_scores.device
        B, A, nc = pd_scores.shape
        R = pd_bboxes.shape[-1]
        G = gt_bboxes.size(1)
        self.bs, self.n_max_boxes = B, G
        if G == 0:
            z = torch.zeros_like(pd_scores[..., 0])
            return (torch.full_like(z, self.bg_idx, dtype=torch.int64), torch.zeros(B, A, 4, device=dev),
                    torch.zeros_like(pd_scores), torch.zeros(B, 0, A, dtype=torch.bool, device=dev),
                    torch.zeros(B, A, dtype=torch.int64, device=dev), torch.zeros(0, R, device=dev),
                    torch.zeros(0, device=dev), torch.zeros(B, A, dtype=torch.bool, device=dev))
        if grid is not None:
            shapes, strides = grid
        elif ss is not None and imgsz is not None:
            shapes, strides = self._grid(ss, imgsz, A)
        else:
            shapes, strides = self._grid_from_anchors(anc_points, stride_tensor, A)
        cgrid = L.make_grid(shapes, strides)
        scores = pd_scores.float().contiguous()
        rays = pd_bboxes.float().contiguous()
        pv = L.PredView()
        off = 0
        for l, (h, w) in enumerate(shapes):
            pv.rays[l] = rays.data_ptr() + off * R * 4
            pv.cls[l] = scores.data_ptr() + off * nc * 4
            pv.rays_sb[l], pv.rays_sa[l], pv.rays_sc[l] = A * R, R, 1
            pv.cls_sb[l], pv.cls_sa[l], pv.cls_sc[l] = A * nc, nc, 1
            pv.ray_scale[l] = 1.0
            off += h * w
        pv.cls_is_logit = 0
        gt, keep = gt_struct(gt_labels, gt_bboxes, gt_coor, mask_gt)
        boxes_h = gt_bboxes.detach().float().cpu().contiguous()
        lib = L.lib()
        cap = int(lib.ycr_candidate_bound_h(C.byref(cgrid), boxes_h.data_ptr(), 4, B * G)) + 64
        cfg = L.AssignCfg(int(self.topk), int(nc), int(R), float(self.alpha), float(self.beta), float(self.eps))
        nbytes = lib.ycr_assign_workspace_bytes(C.byref(cgrid), B, G, C.byref(cfg), cap)
        if nbytes == 0:
            L.check(-1, "ycr_assign_workspace_bytes")
        ws = L.Workspace.get("assign", nbytes, dev)
        pos_cap = B * G * int(self.topk)
        o = {
            "labels": torch.empty(B, A, dtype=torch.int64, device=dev),
            "bboxes": torch.empty(B, A, 4, device=dev),
            "scores": torch.empty(B, A, nc, device=dev),
            "mask_pos": torch.empty(B, G, A, dtype=torch.bool, device=dev),
            "tgi": torch.empty(B, A, dtype=torch.int64, device=dev),
            "fg": torch.empty(B, A, dtype=torch.bool, device=dev),
            "dist": torch.empty(pos_cap, R, device=dev),
            "cent": torch.empty(pos_cap, device=dev),
            "npos": torch.zeros(1, dtype=torch.int32, device=dev),
        }
        out = L.AssignOut()
        out.target_labels_i64 = o["labels"].data_ptr()
        out.target_bboxes = o["bboxes"].data_ptr()
        out.target_scores = o["scores"].data_ptr()
        out.mask_pos = o["mask_pos"].data_ptr()
        out.target_gt_idx_i64 = o["tgi"].data_ptr()
        out.fg_mask = o["fg"].data_ptr()
        out.gt_dist = o["dist"].data_ptr()
        out.centerness = o["cent"].data_ptr()
        out.pos_capacity = pos_cap
        out.n_pos_d = o["npos"].data_ptr()
        if self.debug_metrics:
            o["overlaps"] = torch.empty(B, G, A, device=dev)
            o["align"] = torch.empty(B, G, A, device=dev)
            out.overlaps, out.align_metric = o["overlaps"].data_ptr(), o["align"].data_ptr()
        rc = lib.ycr_assign(C.byref(cgrid), C.byref(pv), C.byref(gt), C.byref(cfg), C.byref(out), ws.data_ptr(),
                            ws.numel(), cap, L.stream_ptr(dev))
        L.check(rc, "ycr_assign")
        P = int(o["npos"].item())
        if P < 0:
            raise L.YcrError("ycr_assign: candidate capacity exceeded (workspace sized for too few candidates)")
        if self.debug_metrics:
            self.last_overlaps, self.last_align_metric = o["overlaps"], o["align"]
        del keep
        return (o["labels"], o["bboxes"], o["scores"], o["mask_pos"], o["tgi"], o["dist"][:P], o["cent"][:P], o["fg"])


def MaskIOU(target, pred):
    """utils/tal.py:1445-1464 — Polar-IoU of (N,R) ray sets; tiny elementwise helper kept for API parity
    (the kernels fuse it into the candidate pass)."""
    both = torch.stack([pred, target], -1)
    l_max = both.max(dim=-1)[0]
    l_min = both.min(dim=-1)[0].clamp(min=1e-6)
    return l_min.sum(dim=-1) / l_max.sum(dim=-1)
