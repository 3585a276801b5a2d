"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/ycr_b200.h
declares; host-only entry points behave; the product refuses to run without CUDA (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest
import torch

from util import ROOT
import ycr_b200  # noqa: F401
from ycr_b200 import _lib as L
from ycr_b200 import build as B


@pytest.fixture(scope="module")
def lib():
    B.build()
    return L.lib()


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "ycr_b200.h")).read()
    declared = set(re.findall(r"\b(ycr_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    raw = C.CDLL(L.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), f"{name} declared in the header but not exported"
    assert declared == set(L.EXPORTS), "ctypes binding and header disagree"
    assert lib.ycr_version() == 100


def test_struct_mirrors_match_the_library(lib):
    """The ctypes mirrors of the seven ABI structs have the sizes the compiled library uses."""
    sizes = (C.c_int * 7)()
    assert lib.ycr_abi_sizes(sizes) == 7
    mirrors = [L.Grid, L.PredView, L.Gt, L.AssignCfg, L.AssignOut, L.LossCfg, L.NmsCfg]
    assert list(sizes) == [C.sizeof(m) for m in mirrors]


def test_host_only_entry_points(lib):
    g = L.make_grid([(80, 80), (40, 40), (20, 20)], [8, 16, 32])
    boxes = torch.tensor([[100., 100., 260., 200.], [0., 0., 0., 0.], [10., 10., 630., 630.]])
    n = lib.ycr_candidate_bound_h(C.byref(g), boxes.data_ptr(), 4, 3)
    exact = 0
    for x1, y1, x2, y2 in boxes.tolist():
        for (h, w), s in zip([(80, 80), (40, 40), (20, 20)], [8, 16, 32]):
            xs = [(i + 0.5) * s for i in range(w)]
            ys = [(i + 0.5) * s for i in range(h)]
            exact += sum(x1 < x < x2 for x in xs) * sum(y1 < y < y2 for y in ys)
    assert exact <= n <= exact * 1.6 + 64
    cfg = L.AssignCfg(10, 80, 36, 0.5, 4.0, 1e-9)
    a = lib.ycr_assign_workspace_bytes(C.byref(g), 2, 8, C.byref(cfg), 10000)
    b = lib.ycr_seg_loss_workspace_bytes(C.byref(g), 2, 8, C.byref(cfg), 10000)
    assert 0 < a < b < 64 << 20
    bad = L.AssignCfg(10, 80, 35, 0.5, 4.0, 1e-9)
    assert lib.ycr_assign_workspace_bytes(C.byref(g), 2, 8, C.byref(bad), 10000) == 0
    assert b"rays" in lib.ycr_last_error()
    ncfg = L.NmsCfg(0.25, 0.7, 0, 0, 300, 80, 30000, 7680.0, None, 0)
    assert lib.ycr_nms_workspace_bytes(4, 8400, 192, C.byref(ncfg)) > 0


def test_stage_targets_host_half_matches_preprocess(lib):
    """ycr_stage_targets_h is the host half of GT packing (no CUDA call): rows, G, the (image, slot) -> row table and the
    candidate bound of a ragged batch with an empty image; applying the table the way the packing kernel does gives the
    oracle's preprocess (utils/loss.py:215-239) output."""
    import numpy as np
    from ycr_b200 import synth
    from oracle import polar_oracle as po
    cfg = synth.PathConfig("st", 4, 5, 160, nc=7)
    batch = synth.make_gts(cfg, 3, ragged=True)                 # image 1 has no GT
    bi, cls, bb = batch["batch_idx"].clone(), batch["cls"].view(-1).clone(), batch["bboxes"].clone()
    segs = [s.contiguous() for s in batch["segments"]]
    N, B = bi.numel(), cfg.batch
    g = L.make_grid(cfg.level_shapes, list(cfg.strides))
    stage = torch.full((N * 726 + B * N,), -7.0)
    ptrs = (C.c_void_p * len(segs))(*[t.data_ptr() for t in segs])
    nrow = (C.c_int * len(segs))(*[t.shape[0] for t in segs])
    G, cap = C.c_int(0), C.c_int64(0)
    rc = lib.ycr_stage_targets_h(bi.data_ptr(), cls.data_ptr(), bb.data_ptr(), ptrs, nrow, len(segs), N, B, C.byref(g),
                                 C.c_float(160.0), C.c_float(160.0), stage.data_ptr(), C.byref(G), C.byref(cap))
    assert rc == 0, lib.ycr_last_error()
    counts = torch.bincount(bi.long(), minlength=B)
    assert G.value == int(counts.max()) and int(counts[1]) == 0
    head = stage[:N * 6].view(N, 6)
    seg = stage[N * 6:N * 726].view(N, 720)
    assert torch.equal(head[:, 0], bi) and torch.equal(head[:, 1], cls) and torch.equal(head[:, 2:], bb)
    assert torch.equal(seg, torch.cat([t.reshape(-1, 720) for t in segs]))
    row_of = np.frombuffer(stage.numpy().tobytes(), dtype=np.int32)[N * 726:N * 726 + B * G.value].reshape(B, G.value)
    # what k_pack_targets_mapped does with the table
    packed = torch.zeros(B, G.value, 725)
    scale = torch.tensor([160.0, 160.0] * 360)
    for b in range(B):
        for sl in range(G.value):
            n = int(row_of[b, sl])
            assert (n >= 0) == (sl < int(counts[b]))
            if n >= 0:
                assert int(bi[n]) == b
                x, y, w, h = (head[n, 2:6] * 160.0).tolist()
                packed[b, sl, 0] = head[n, 1]
                packed[b, sl, 1:5] = torch.tensor([x - w / 2, y - h / 2, x + w / 2, y + h / 2])
                packed[b, sl, 5:] = seg[n] * scale
    ref = po.pack_targets(batch, B, (160, 160))
    assert ref.shape == packed.shape
    assert torch.equal(packed[..., 0], ref[..., 0]) and torch.equal(packed[..., 5:], ref[..., 5:])
    assert float((packed[..., 1:5] - ref[..., 1:5]).abs().max()) < 1e-3
    # the bound covers the exact number of in-box anchors
    exact = int(po.in_box_mask(po.make_anchors(cfg.level_shapes, cfg.strides)[0] * po.make_anchors(cfg.level_shapes, cfg.strides)[1],
                               ref[..., 1:5]).sum())
    assert exact <= cap.value
    # malformed input: the contour blocks do not add up to the boxes
    nrow_bad = (C.c_int * len(segs))(*[max(0, t.shape[0] - 1) for t in segs])
    rc = lib.ycr_stage_targets_h(bi.data_ptr(), cls.data_ptr(), bb.data_ptr(), ptrs, nrow_bad, len(segs), N, B, C.byref(g),
                                 C.c_float(160.0), C.c_float(160.0), stage.data_ptr(), C.byref(G), C.byref(cap))
    assert rc == -1 and b"segment rows" in lib.ycr_last_error()


def test_argument_errors_without_gpu(lib):
    g = L.make_grid([(20, 20)], [8])
    rc = lib.ycr_decode(C.byref(g), None, 1, 10, 36, None, None)
    assert rc == -1 and b"null" in lib.ycr_last_error()
    ncfg = L.NmsCfg(1.5, 0.7, 0, 0, 300, 80, 30000, 7680.0, None, 0)
    one = torch.zeros(8)
    rc = lib.ycr_nms(one.data_ptr(), 1, 192, 100, C.byref(ncfg), one.data_ptr(), one.data_ptr(), one.data_ptr(), 8, None)
    assert rc == -1 and b"thresholds" in lib.ycr_last_error()


def test_no_cpu_fallback():
    from ycr_b200.head import decode
    from ycr_b200.ops import non_max_suppression
    from ycr_b200.loss import v8SegmentationLoss
    with pytest.raises(L.YcrError):
        decode([torch.zeros(1, 46, 4, 4)], [8], 10, 36)
    with pytest.raises(L.YcrError):
        non_max_suppression(torch.zeros(1, 122, 16), nc=10)
    crit = v8SegmentationLoss(nc=10, nm=36, strides=(8, 16, 32), device="cpu")
    with pytest.raises(L.YcrError):
        crit(([torch.zeros(1, 46, 4, 4)] * 3, 5, 2), {})


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "yolo-contour-regression_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|polar_oracle|oracle/", src, re.M), \
                    f"{f} references the oracle"
