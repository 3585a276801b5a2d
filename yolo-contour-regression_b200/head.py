"""Host-side mirror of the polar `Segment` head (nn/modules/head.py:436-574 with the parent that
matches it, polarpaperDetect nn/modules/head.py:364-433; paths relative to
/root/reference/ultralytics-main/ultralytics/).  The convolutions stay in cuDNN (north_star); the
eval-branch decode (distance2mask, nn/modules/head.py:461-494) is one CUDA kernel."""
from __future__ import annotations

import ctypes as C
import math

import torch
import torch.nn as nn

from . import _lib as L


def decode(feats, strides, nc: int, R: int = 36):
    """feats: list of (B, R+nc, H_l, W_l) -> allpred (B, 4+nc+3R, A), as distance2mask returns it."""
    L.require_cuda(*feats)
    dt = feats[0].dtype if feats[0].dtype in L.DTYPE_CODE else torch.float32   # fp16 / bf16 maps are read in place
    feats = [f if (f.dtype == dt and f.is_contiguous()) else f.to(dt).contiguous() for f in feats]
    B = feats[0].shape[0]
    if feats[0].shape[1] != R + nc:
        raise ValueError(f"feature maps have {feats[0].shape[1]} channels, expected {R + nc}")
    shapes = [tuple(f.shape[2:]) for f in feats]
    A = sum(h * w for h, w in shapes)
    out = torch.empty(B, 4 + nc + 3 * R, A, device=feats[0].device, dtype=torch.float32)
    # per anchor {best class score, best class}: a by-product of the class pass that single-label NMS starts
    # from.  It rides on the output tensor; ops.non_max_suppression uses it only if the tensor it is handed is
    # this very object and no in-place operation has touched it since (version counter).
    best = torch.empty(B, A, 2, device=feats[0].device, dtype=torch.int32)
    if B == 0:
        return out
    cgrid = L.make_grid(shapes, [float(s) for s in strides])
    rc = L.lib().ycr_decode_dt(C.byref(cgrid), L.ptr_array(feats), L.DTYPE_CODE[dt], B, nc, R, out.data_ptr(),
                               best.data_ptr(), L.stream_ptr(out.device))
    L.check(rc, "ycr_decode_dt")
    out._ycr_best_class = (best, out._version, nc)
    out._ycr_feats = (feats, cgrid, L.DTYPE_CODE[dt], R)   # lets NMS rebuild its kept rows from R ray values each
    return out


def _default_act():
    """The activation the model definition selected: parse_model sets `Conv.default_act` from the yaml's `activation:`
    key (nn/tasks.py:668-671; the reference's own yolov8-seg.yaml asks for nn.ReLU()).  SiLU, the class default of
    nn/modules/conv.py, when no `ultralytics` is loaded."""
    import copy
    import sys
    conv_mod = sys.modules.get("ultralytics.nn.modules.conv")
    act = getattr(getattr(conv_mod, "Conv", None), "default_act", None) if conv_mod is not None else None
    return copy.deepcopy(act) if isinstance(act, nn.Module) else nn.SiLU()


class _ConvBnSiLU(nn.Module):
    """Conv2d + BatchNorm2d + activation (SiLU by default), the block the head stacks (nn/modules/conv.py Conv)."""

    def __init__(self, c1, c2, k=3):
        super().__init__()
        self.conv = nn.Conv2d(c1, c2, k, 1, k // 2, bias=False)
        self.bn = nn.BatchNorm2d(c2)
        self.act = _default_act()

    def forward(self, x):
        return self.act(self.bn(self.conv(x)))


class _NoProto(int):
    """The third element of the eval output.  The reference returns the int 1 there (nn/modules/head.py:570), and its
    own predictor then indexes it as a proto tensor (models/yolo/segment/predict.py:26,39: `proto[i]`), which raises
    as soon as an image has a detection.  This is still the int 1, but it can be indexed."""

    def __getitem__(self, i):
        return self


class Segment(nn.Module):
    """`Segment(nc=80, nm=36, npr=256, ch=())` — nn/modules/head.py:439.

    forward(x: list[Tensor]):
      training -> (feats, 5, 2)                     nn/modules/head.py:555-558
      eval     -> (allpred, (feats, allpred, 1))    nn/modules/head.py:559-570
      export   -> (rays, cls logits)                nn/modules/head.py:572-574
    feats[l] is (B, nm+nc, H_l, W_l): nm ray channels then nc class logits (polarpaperDetect.forward,
    nn/modules/head.py:388-392)."""
    dynamic = False
    export = False
    shape = None

    def __init__(self, nc=80, nm=36, npr=256, ch=()):
        super().__init__()
        self.nc = nc
        self.nm = nm
        self.npr = npr
        self.nl = len(ch)
        self.reg_max = 16
        self.no = nc + nm
        self.stride = torch.zeros(self.nl)
        self.anchors = torch.empty(0)   # BaseModel._apply moves stride / anchors / strides of the head (nn/tasks.py:184-187)
        self.strides = torch.empty(0)
        c2, c3 = max((16, ch[0] // 4, self.reg_max * 4)), max(ch[0], min(self.nc, 100))
        self.cv2 = nn.ModuleList(nn.Sequential(_ConvBnSiLU(x, c2, 3), _ConvBnSiLU(c2, c2, 3), nn.Conv2d(c2, self.nm, 1))
                                 for x in ch)
        self.cv3 = nn.ModuleList(nn.Sequential(_ConvBnSiLU(x, c3, 3), _ConvBnSiLU(c3, c3, 3), nn.Conv2d(c3, self.nc, 1))
                                 for x in ch)

    def bias_init(self):
        """nn/modules/head.py:427-433"""
        for a, b, s in zip(self.cv2, self.cv3, self.stride):
            a[-1].bias.data[:] = 1.0
            b[-1].bias.data[:self.nc] = math.log(5 / self.nc / (640 / float(s)) ** 2)

    def forward(self, x):
        feats = [torch.cat((self.cv2[i](x[i]), self.cv3[i](x[i])), 1) for i in range(self.nl)]
        if self.export:
            cat = torch.cat([f.view(f.shape[0], self.no, -1) for f in feats], 2)
            return cat[:, :self.nm], cat[:, self.nm:]
        if self.training:
            return feats, 5, 2
        allpred = decode(feats, self.stride.tolist(), self.nc, self.nm)
        return allpred, (feats, allpred, _NoProto(1))
