"""ctypes binding of the C-ABI library (include/ycr_b200.h).  No fallback: if the library is
missing or a call fails, the caller gets an exception."""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libycr_b200.so")
MAX_LEVELS = 4
CONTOUR_POINTS = 360


class Grid(C.Structure):
    _fields_ = [("n_levels", C.c_int), ("h", C.c_int * MAX_LEVELS), ("w", C.c_int * MAX_LEVELS),
                ("stride", C.c_float * MAX_LEVELS)]


class PredView(C.Structure):
    _fields_ = [("rays", C.c_void_p * MAX_LEVELS), ("cls", C.c_void_p * MAX_LEVELS),
                ("rays_sb", C.c_int64 * MAX_LEVELS), ("rays_sa", C.c_int64 * MAX_LEVELS),
                ("rays_sc", C.c_int64 * MAX_LEVELS),
                ("cls_sb", C.c_int64 * MAX_LEVELS), ("cls_sa", C.c_int64 * MAX_LEVELS),
                ("cls_sc", C.c_int64 * MAX_LEVELS),
                ("ray_scale", C.c_float * MAX_LEVELS), ("cls_is_logit", C.c_int)]


class Gt(C.Structure):
    _fields_ = [("B", C.c_int), ("G", C.c_int),
                ("labels", C.c_void_p), ("labels_stride", C.c_int64),
                ("boxes", C.c_void_p), ("boxes_stride", C.c_int64),
                ("coor", C.c_void_p), ("coor_stride", C.c_int64),
                ("mask_gt", C.c_void_p), ("mask_stride", C.c_int64)]


class AssignCfg(C.Structure):
    _fields_ = [("topk", C.c_int), ("num_classes", C.c_int), ("rays", C.c_int),
                ("alpha", C.c_float), ("beta", C.c_float), ("eps", C.c_float)]


class AssignOut(C.Structure):
    _fields_ = [("target_labels_i64", C.c_void_p), ("target_bboxes", C.c_void_p),
                ("target_scores", C.c_void_p), ("mask_pos", C.c_void_p),
                ("target_gt_idx_i64", C.c_void_p), ("fg_mask", C.c_void_p),
                ("gt_dist", C.c_void_p), ("centerness", C.c_void_p), ("pos_capacity", C.c_int),
                ("n_pos_d", C.c_void_p), ("overlaps", C.c_void_p), ("align_metric", C.c_void_p)]


class LossCfg(C.Structure):
    _fields_ = [("box_gain", C.c_float), ("cls_gain", C.c_float)]


class NmsCfg(C.Structure):
    _fields_ = [("conf_thres", C.c_float), ("iou_thres", C.c_float), ("agnostic", C.c_int),
                ("multi_label", C.c_int), ("max_det", C.c_int), ("nc", C.c_int), ("max_nms", C.c_int),
                ("max_wh", C.c_float), ("classes", C.c_void_p), ("n_classes", C.c_int), ("compact_rows", C.c_int),
                ("best_class", C.c_void_p), ("feats", C.c_void_p * MAX_LEVELS), ("grid", C.POINTER(Grid)),
                ("feats_dtype", C.c_int), ("rays", C.c_int)]


EXPORTS = {
    "ycr_last_error": (C.c_char_p, []),
    "ycr_version": (C.c_int, []),
    "ycr_abi_sizes": (C.c_int, [C.c_void_p]),
    "ycr_profile_begin": (C.c_int, [C.c_int]),
    "ycr_profile_select": (C.c_int, [C.c_uint]),
    "ycr_profile_end": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ycr_debug_stats": (C.c_int, [C.c_void_p, C.c_int]),
    "ycr_candidate_bound_h": (C.c_int64, [C.POINTER(Grid), C.c_void_p, C.c_int64, C.c_int]),
    "ycr_assign_workspace_bytes": (C.c_size_t, [C.POINTER(Grid), C.c_int, C.c_int, C.POINTER(AssignCfg), C.c_int64]),
    "ycr_assign": (C.c_int, [C.POINTER(Grid), C.POINTER(PredView), C.POINTER(Gt), C.POINTER(AssignCfg),
                             C.POINTER(AssignOut), C.c_void_p, C.c_size_t, C.c_int64, C.c_void_p]),
    "ycr_seg_loss_workspace_bytes": (C.c_size_t, [C.POINTER(Grid), C.c_int, C.c_int, C.POINTER(AssignCfg), C.c_int64]),
    "ycr_seg_loss_fwd_bwd": (C.c_int, [C.POINTER(Grid), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                       C.POINTER(Gt), C.POINTER(AssignCfg), C.POINTER(LossCfg), C.c_void_p,
                                       C.c_void_p, C.c_size_t, C.c_int64, C.c_void_p]),
    "ycr_seg_loss_fwd_bwd_dt": (C.c_int, [C.POINTER(Grid), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int,
                                          C.POINTER(Gt), C.POINTER(AssignCfg), C.POINTER(LossCfg), C.c_void_p,
                                          C.c_void_p, C.c_size_t, C.c_int64, C.c_void_p]),
    "ycr_scale_grads_dt": (C.c_int, [C.POINTER(Grid), C.c_int, C.c_int, C.POINTER(C.c_void_p), C.c_int, C.c_void_p,
                                     C.c_void_p]),
    "ycr_scale_grads": (C.c_int, [C.POINTER(Grid), C.c_int, C.c_int, C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p]),
    "ycr_pack_targets": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float,
                                   C.c_void_p, C.c_void_p]),
    "ycr_pack_targets_split": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_float,
                                         C.c_float, C.c_void_p, C.c_void_p]),
    "ycr_candidate_bound_xywhn_h": (C.c_int64, [C.POINTER(Grid), C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_float]),
    "ycr_stage_targets_h": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.c_int,
                                      C.c_int, C.c_int, C.POINTER(Grid), C.c_float, C.c_float, C.c_void_p,
                                      C.POINTER(C.c_int), C.POINTER(C.c_int64)]),
    "ycr_pack_targets_mapped": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_float,
                                          C.c_float, C.c_void_p, C.c_void_p]),
    "ycr_resample_segments": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "ycr_bbox_loss_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "ycr_bbox_loss_fwd_bwd": (C.c_int, [C.c_void_p] * 7 + [C.c_int] * 5 + [C.c_void_p] * 4 + [C.c_size_t, C.c_void_p]),
    "ycr_decode": (C.c_int, [C.POINTER(Grid), C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_void_p,
                             C.c_void_p]),
    "ycr_decode_best": (C.c_int, [C.POINTER(Grid), C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_void_p,
                                  C.c_void_p, C.c_void_p]),
    "ycr_detect_workspace_bytes": (C.c_size_t, [C.POINTER(Grid), C.c_int, C.POINTER(NmsCfg)]),
    "ycr_detect": (C.c_int, [C.POINTER(Grid), C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(NmsCfg),
                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "ycr_rasterize_contours": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "ycr_mask_iou_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int64]),
    "ycr_mask_iou": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_float, C.c_void_p,
                               C.c_void_p, C.c_size_t, C.c_void_p]),
    "ycr_decode_dt": (C.c_int, [C.POINTER(Grid), C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                C.c_void_p, C.c_void_p]),
    "ycr_nms_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.POINTER(NmsCfg)]),
    "ycr_nms": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(NmsCfg), C.c_void_p, C.c_void_p,
                          C.c_void_p, C.c_size_t, C.c_void_p]),
}

_lib = None


class YcrError(RuntimeError):
    pass


def lib():
    """The loaded library.  Raises if it has not been built (python -c 'import __graft_entry__ as g; g.build()')."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise YcrError(f"{LIB_PATH} not found: the CUDA library is not built and there is no CPU fallback")
        h = C.CDLL(LIB_PATH)
        for name, (res, args) in EXPORTS.items():
            fn = getattr(h, name)
            fn.restype = res
            fn.argtypes = args
        _lib = h
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().ycr_last_error().decode()
        if rc == -1:
            raise ValueError(f"{what}: {msg}")
        raise YcrError(f"{what} failed ({rc}): {msg}")


DTYPE_CODE = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}


def stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def make_grid(level_shapes, strides) -> Grid:
    g = Grid()
    g.n_levels = len(level_shapes)
    if g.n_levels > MAX_LEVELS:
        raise ValueError(f"at most {MAX_LEVELS} levels supported")
    for i, ((h, w), s) in enumerate(zip(level_shapes, strides)):
        g.h[i], g.w[i], g.stride[i] = int(h), int(w), float(s)
    return g


def ptr_array(tensors):
    arr = (C.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr() if t is not None else None
    return arr


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise YcrError("ycr_b200 kernels need CUDA tensors; there is no CPU path")


class Workspace:
    """Grow-only device scratch buffer, one per (purpose, device, stream): calls on different streams must not
    share scratch memory, calls on one stream are ordered."""
    _bufs: dict = {}

    @classmethod
    def get(cls, key, nbytes, device):
        k = (key, str(device), torch.cuda.current_stream(device).cuda_stream)
        buf = cls._bufs.get(k)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(int(nbytes * 1.25) + 1024, dtype=torch.uint8, device=device)
            cls._bufs[k] = buf
        return buf
