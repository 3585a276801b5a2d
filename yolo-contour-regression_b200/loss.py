"""Host-side mirror of `ultralytics/utils/loss.py` for the polar path (paths relative to
/root/reference/ultralytics-main/ultralytics/): `v8SegmentationLoss` (utils/loss.py:772-878) and
`MaskIOULoss` (utils/loss.py:109-127).  One C-ABI call runs assignment + polar targets + both loss
terms + the gradient with respect to the head outputs."""
from __future__ import annotations

import ctypes as C
from types import SimpleNamespace

import torch
import torch.nn as nn

from . import _lib as L
from .tal import TaskAlignedAssigner, gt_struct


class MaskIOULoss(nn.Module):
    """utils/loss.py:109-127 — kept for API parity (`criterion.polar_loss`); the fused kernel computes
    the same quantity for the positives without materialising them."""

    def forward(self, pred_rays, target_rays, target_scores, target_scores_sum):
        weight = target_scores.sum(-1)
        total = torch.stack([pred_rays, target_rays], -1)
        l_max = total.max(dim=2)[0]
        l_min = total.min(dim=2)[0].clamp(min=1e-6)
        loss = (l_max.sum(dim=1) / l_min.sum(dim=1)).log() * weight
        return loss.sum() / target_scores_sum


class _SegLossFn(torch.autograd.Function):
    """loss = f(feat_0, feat_1, feat_2); forward already produced d loss / d feat_l."""

    @staticmethod
    def forward(ctx, crit, gt, cand_cap, *feats):
        dev = feats[0].device
        lib = L.lib()
        B = feats[0].shape[0]
        shapes = [tuple(f.shape[2:]) for f in feats]
        cgrid = L.make_grid(shapes, crit.stride_list)
        need_grad = any(f.requires_grad for f in feats)
        grads = [torch.empty_like(f) for f in feats] if need_grad else None
        loss_out = torch.empty(4, device=dev, dtype=torch.float32)
        nbytes = lib.ycr_seg_loss_workspace_bytes(C.byref(cgrid), B, gt.G, C.byref(crit.acfg), cand_cap)
        if nbytes == 0:
            L.check(-1, "ycr_seg_loss_workspace_bytes")
        ws = L.Workspace.get("seg_loss", nbytes, dev)
        fp = L.ptr_array(feats)
        gp = L.ptr_array(grads) if grads is not None else None
        dt = L.DTYPE_CODE[feats[0].dtype]
        rc = lib.ycr_seg_loss_fwd_bwd_dt(C.byref(cgrid), fp, gp, dt, C.byref(gt), C.byref(crit.acfg), C.byref(crit.lcfg),
                                         loss_out.data_ptr(), ws.data_ptr(), ws.numel(), cand_cap, L.stream_ptr(dev))
        L.check(rc, "ycr_seg_loss_fwd_bwd_dt")
        ctx.dt = dt
        ctx.grads = grads
        ctx.scaled = False
        ctx.cgrid = cgrid
        ctx.channels = feats[0].shape[1]
        ctx.B = B
        ctx.mark_non_differentiable(loss_out)
        # the total shares loss_out's storage (detached alias, not a tracked view: the trainer's in-place
        # `loss *= world_size`, engine/trainer.py:365, stays legal) - a clone would be one more copy on the stream
        return loss_out[0].detach(), loss_out

    @staticmethod
    def backward(ctx, g_total, _g_items):
        if ctx.scaled:
            # a second backward through the same graph (retain_graph=True): the maps were handed to autograd by the
            # first one (and carry its upstream gradient)
            raise RuntimeError("v8SegmentationLoss: backward through the same loss twice is not supported "
                               "(the gradient maps are produced once, by the forward kernel)")
        grads = ctx.grads
        if grads is None:
            return (None, None, None) + (None,) * 3
        # hand the maps over: while this node kept a reference, AccumulateGrad could not adopt a map as the .grad of a
        # leaf input and copied it instead (three device-to-device copies, 250 MB at C2, a tenth of the step)
        ctx.grads = None
        dev = grads[0].device
        g = g_total.detach().to(device=dev, dtype=torch.float32).contiguous()
        ctx.scaled = True
        rc = L.lib().ycr_scale_grads_dt(C.byref(ctx.cgrid), ctx.B, ctx.channels, L.ptr_array(grads), ctx.dt, g.data_ptr(),
                                        L.stream_ptr(dev))
        L.check(rc, "ycr_scale_grads_dt")
        return (None, None, None) + tuple(grads)


class v8SegmentationLoss:
    """utils/loss.py:772-878 (+ base v8DetectionLoss.__init__ utils/loss.py:192-214).

    `__call__(preds, batch)` -> `(loss.sum() * batch_size  [with grad], loss.detach() (2,))`, where
    loss = [box_gain * polar_iou_loss, cls_gain * bce] — identical to the reference's return."""

    def __init__(self, model=None, *, nc=None, nm=36, strides=None, box=7.5, cls=0.5, device=None, global_norm=False):
        if model is not None:
            device = next(model.parameters()).device
            h = model.args
            m = model.model[-1]
            self.hyp = h
            self.stride = m.stride
            self.nc = m.nc
            self.no = m.no
            self.reg_max = getattr(m, "reg_max", 16)
            self.nm = m.nm
            self.overlap = getattr(h, "overlap_mask", True)
            box, cls = float(h.box), float(h.cls)
        else:
            self.hyp = SimpleNamespace(box=box, cls=cls)
            self.stride = torch.tensor(strides, dtype=torch.float32)
            self.nc, self.nm = nc, nm
            self.no = nc + nm
            self.reg_max = 16
            self.overlap = True
        self.device = torch.device(device)
        # False = the reference's DDP semantics (local normaliser); True = one scalar all-reduce per step (dp.py)
        self.global_norm = bool(global_norm)
        self.stride_list = [float(s) for s in (self.stride.tolist() if torch.is_tensor(self.stride) else self.stride)]
        self.use_dfl = self.reg_max > 1
        self.assigner = TaskAlignedAssigner(topk=10, num_classes=self.nc, alpha=0.5, beta=4.0)  # utils/loss.py:210
        self.polar_loss = MaskIOULoss()
        self.rays = int(self.nm)  # the reference keeps the first 36 of nm channels (utils/loss.py:818); here nm == rays
        self.acfg = L.AssignCfg(10, int(self.nc), int(self.rays), 0.5, 4.0, 1e-9)
        self.lcfg = L.LossCfg(float(box), float(cls))

    # -- GT packing: utils/loss.py:834-844 + preprocess utils/loss.py:215-239 ----------------------
    def _staging(self, n_rows, extra=0):
        """One of two pinned host buffers for n_rows GT rows (6 header floats + 720 contour floats per row, header
        block first), used in turn: the host only waits for the copy that read THIS buffer two steps ago."""
        if getattr(self, "_stage_bufs", None) is None:
            self._stage_bufs = [None, None]
            self._stage_evs = [None, None]
            self._stage_turn = 0
        k = self._stage_turn = 1 - self._stage_turn
        if self._stage_evs[k] is not None:
            self._stage_evs[k].synchronize()
            self._stage_evs[k] = None
        buf = self._stage_bufs[k]
        need = n_rows * 726 + extra
        if buf is None or buf.numel() < need:
            buf = self._stage_bufs[k] = torch.empty(max(need, 64 * 726), dtype=torch.float32).pin_memory()
        return buf[:need]

    @property
    def last_gt_copy_event(self):
        """CUDA event recorded right behind the latest host-to-device copy of GT rows (None if there was none
        or it has been waited for).  A caller that prefetches the next batch on another stream makes that stream
        wait for it, so the big copy does not overtake the small one this step's kernels depend on."""
        return getattr(self, "_stage_ev", None)

    def pack_targets(self, batch, batch_size, img_hw):
        """-> (packed (B,G,725) device tensor, candidate upper bound).  The rows arrive on the host from the
        dataloader: `batch['segments']` is concatenated straight into a pinned staging buffer (the one host pass over
        the contours), the six header columns next to it, one H2D copy, and a kernel pads and scales them."""
        dev = self.device
        bi = batch["batch_idx"].view(-1)
        N = bi.numel()
        h, w = float(img_hw[0]), float(img_hw[1])
        if N == 0:
            return torch.zeros(batch_size, 0, 5 + 720, device=dev), 0
        segs = batch["segments"]
        lib = L.lib()
        cgrid = L.make_grid(self._shapes, self.stride_list)
        seg_list = list(segs) if isinstance(segs, (list, tuple)) else [segs]
        on_host = all(t.device.type == "cpu" for t in seg_list) and bi.device.type == "cpu"
        mapped = False
        if on_host:
            stage = self._staging(N, batch_size * N)   # rows + the (image, slot) -> row table (at most B * N ints)
            cls_t, box_t = batch["cls"].view(-1), batch["bboxes"].view(-1, 4)
            plain = (bi.dtype == torch.float32 and cls_t.dtype == torch.float32 and box_t.dtype == torch.float32 and
                     box_t.is_contiguous() and all(t.dtype == torch.float32 and t.is_contiguous() for t in seg_list))
            if plain:
                # one C call: memcpy of every image's contours into the pinned buffer, the header columns, G and the
                # candidate bound (the torch ops this replaces cost more host time than the kernels take)
                nt = len(seg_list)
                ptrs = (C.c_void_p * nt)(*[t.data_ptr() for t in seg_list])
                nrow = (C.c_int * nt)(*[t.numel() // 720 for t in seg_list])
                g_out, cap_out = C.c_int(0), C.c_int64(0)
                rc = lib.ycr_stage_targets_h(bi.data_ptr(), cls_t.data_ptr(), box_t.data_ptr(), ptrs, nrow, nt, N, batch_size,
                                             C.byref(cgrid), w, h, stage.data_ptr(), C.byref(g_out), C.byref(cap_out))
                if rc != 0:   # malformed rows: the reference re-wraps these as TypeError (utils/loss.py:850-856)
                    raise RuntimeError(lib.ycr_last_error().decode())
                G, cap = int(g_out.value), int(cap_out.value) + 64
                mapped = True
                stage = stage[:N * 726 + batch_size * G]
            else:
                stage = stage[:N * 726]
                head = stage[:N * 6].view(N, 6)
                seg = stage[N * 6:].view(N, 720)
                seg.copy_(torch.cat([t.reshape(-1, 720) for t in seg_list], 0))
                head[:, 0] = bi
                head[:, 1] = cls_t
                head[:, 2:6] = box_t
                G = int(torch.bincount(bi.long(), minlength=batch_size).max())
                cap = int(lib.ycr_candidate_bound_xywhn_h(C.byref(cgrid), head.data_ptr() + 8, 6, N, w, h)) + 64
            # The copy goes on its own stream into one of two persistent device buffers: the host runs ahead of the
            # device, so the rows of step k+1 cross the bus while the kernels of step k still run.  The compute
            # stream waits for the copy; the copy waits until the kernel that last read this buffer has finished.
            cur = torch.cuda.current_stream(dev)
            if getattr(self, "_copy_stream", None) is None:
                self._copy_stream = torch.cuda.Stream(dev)
                self._rows_dev = [None, None]
                self._rows_free = [None, None]
                self._rows_turn = 0
            k = self._rows_turn = 1 - self._rows_turn
            if self._rows_dev[k] is None or self._rows_dev[k].numel() < stage.numel():
                self._rows_dev[k] = torch.empty(max(stage.numel(), 64 * 726), device=dev, dtype=torch.float32)
                # The allocator hands out blocks whose previous owner may still be in use by kernels queued on the
                # compute stream (it orders reuse on THAT stream only): the copy stream must not write the new buffer
                # before everything enqueued so far has run.
                ev = torch.cuda.Event()
                ev.record(cur)
                self._rows_free[k] = ev
            rows = self._rows_dev[k][:stage.numel()]
            cs = self._copy_stream
            if self._rows_free[k] is not None:
                cs.wait_event(self._rows_free[k])
            with torch.cuda.stream(cs):
                rows.copy_(stage, non_blocking=True)
            self._stage_ev = torch.cuda.Event()
            self._stage_ev.record(cs)
            self._stage_evs[self._stage_turn] = self._stage_ev
            cur.wait_event(self._stage_ev)
            self._rows_slot = k
        else:
            seg = torch.cat([t.reshape(-1, 720) for t in seg_list], 0).float().to(dev)
            head = torch.cat((bi.view(-1, 1).float().to(dev), batch["cls"].view(-1, 1).float().to(dev),
                              batch["bboxes"].view(-1, 4).float().to(dev)), 1)
            G = int(torch.bincount(bi.long().cpu(), minlength=batch_size).max())
            bb_h = head[:, 2:6].cpu().contiguous()
            cap = int(lib.ycr_candidate_bound_xywhn_h(C.byref(cgrid), bb_h.data_ptr(), 4, N, w, h)) + 64
            rows = torch.cat((head.reshape(-1), seg.reshape(-1)))
        out = torch.empty(batch_size, G, 5 + 720, device=dev)
        if mapped:
            rc = lib.ycr_pack_targets_mapped(rows.data_ptr(), 6, rows.data_ptr() + N * 6 * 4, 720,
                                             rows.data_ptr() + N * 726 * 4, batch_size, G, w, h, out.data_ptr(),
                                             L.stream_ptr(dev))
        else:
            rc = lib.ycr_pack_targets_split(rows.data_ptr(), 6, rows.data_ptr() + N * 6 * 4, 720, N, batch_size, G, w, h,
                                            out.data_ptr(), L.stream_ptr(dev))
        L.check(rc, "ycr_pack_targets")
        if on_host:   # the staged rows may be overwritten once this kernel has read them
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(dev))
            self._rows_free[self._rows_slot] = ev
        return out, cap

    def __call__(self, preds, batch):
        feats, _, _ = preds if len(preds) == 3 else preds[1]  # utils/loss.py:812
        L.require_cuda(*feats)
        # fp32, fp16 and bf16 maps are read in place (autocast is the reference's default, engine/trainer.py:332)
        dt = feats[0].dtype if feats[0].dtype in L.DTYPE_CODE else torch.float32
        feats = [f if (f.dtype == dt and f.is_contiguous()) else f.to(dt).contiguous() for f in feats]
        B = feats[0].shape[0]
        if feats[0].shape[1] != self.rays + self.nc:
            raise ValueError(f"head emits {feats[0].shape[1]} channels, expected rays+nc = {self.rays + self.nc}")
        self._shapes = [tuple(f.shape[2:]) for f in feats]
        img_hw = (feats[0].shape[2] * self.stride_list[0], feats[0].shape[3] * self.stride_list[0])
        try:
            packed, cap = self.pack_targets(batch, B, img_hw)
        except L.YcrError:   # a missing library or a failed CUDA call is not a dataset problem
            raise
        except RuntimeError as e:  # same re-wrap as utils/loss.py:850-856 (cat / view of malformed rows)
            raise TypeError("ERROR segment dataset incorrectly formatted or not a segment dataset.") from e
        gt_labels, gt_boxes, gt_coor = packed.split((1, 4, 720), 2)
        gt, keep = gt_struct(gt_labels, gt_boxes, gt_coor, None)
        total, out = _SegLossFn.apply(self, gt, cap, *feats)
        del keep
        items = out[1:3].detach()
        if self.global_norm:
            from .dp import rescale_to_global_norm
            total, items = rescale_to_global_norm(total, items, out[3])
        return total, items

    def call_packed(self, feats, packed, cand_cap):
        """Extension: same as __call__ with GTs already packed on the device ((B,G,725), px)."""
        gt_labels, gt_boxes, gt_coor = packed.split((1, 4, 720), 2)
        gt, keep = gt_struct(gt_labels, gt_boxes, gt_coor, None)
        total, out = _SegLossFn.apply(self, gt, cand_cap, *feats)
        del keep
        return total, out[1:3].detach()


class _BboxLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred_dist, pred_bboxes, anchor_points, target_bboxes, target_scores, tss, fg_mask, reg_max, use_dfl):
        dev = pred_bboxes.device
        lib = L.lib()
        B, A = pred_bboxes.shape[:2]
        nc = target_scores.shape[-1]
        pd = pred_dist.float().contiguous()
        pb = pred_bboxes.float().contiguous()
        need = pred_dist.requires_grad or pred_bboxes.requires_grad
        gd = torch.empty_like(pd) if need else None
        gb = torch.empty_like(pb) if need else None
        out = torch.empty(2, device=dev)
        tss_t = torch.as_tensor(tss, dtype=torch.float32, device=dev).reshape(1).contiguous()
        keep = [anchor_points.float().contiguous(), target_bboxes.float().contiguous(),
                target_scores.float().contiguous(), fg_mask.to(torch.uint8).contiguous(), tss_t]
        ws = L.Workspace.get("bbox_loss", lib.ycr_bbox_loss_workspace_bytes(B, A), dev)
        rc = lib.ycr_bbox_loss_fwd_bwd(pd.data_ptr(), pb.data_ptr(), keep[0].data_ptr(), keep[1].data_ptr(),
                                       keep[2].data_ptr(), keep[3].data_ptr(), keep[4].data_ptr(), B, A, nc,
                                       int(reg_max), int(bool(use_dfl)), out.data_ptr(),
                                       gd.data_ptr() if need else None, gb.data_ptr() if need else None,
                                       ws.data_ptr(), ws.numel(), L.stream_ptr(dev))
        L.check(rc, "ycr_bbox_loss_fwd_bwd")
        ctx.gd, ctx.gb = gd, gb
        return out[0].clone(), out[1].clone()

    @staticmethod
    def backward(ctx, g_iou, g_dfl):
        # the kernel stored d(loss_iou)/d(pred_bboxes) and d(loss_dfl)/d(pred_dist); the two terms do not mix
        gd = ctx.gd * g_dfl if ctx.gd is not None else None
        gb = ctx.gb * g_iou if ctx.gb is not None else None
        return gd, gb, None, None, None, None, None, None, None


class BboxLoss(nn.Module):
    """utils/loss.py:53-87 — CIoU + DFL box loss, same constructor and forward signature; returns
    `(loss_iou, loss_dfl)`.  Not called by the live polar v8SegmentationLoss (utils/loss.py:211 constructs it
    only); provided for the north_star's 'DFL/CIoU box terms' and the fork's `ori*` classes."""

    def __init__(self, reg_max, use_dfl=False):
        super().__init__()
        self.reg_max = reg_max
        self.use_dfl = use_dfl

    def forward(self, pred_dist, pred_bboxes, anchor_points, target_bboxes, target_scores, target_scores_sum, fg_mask):
        L.require_cuda(pred_dist, pred_bboxes, target_bboxes, target_scores, fg_mask)
        return _BboxLossFn.apply(pred_dist, pred_bboxes, anchor_points, target_bboxes, target_scores,
                                 target_scores_sum, fg_mask, self.reg_max, self.use_dfl)
