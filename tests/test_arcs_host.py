"""The per-thread logic of the candidate kernel (csrc/polar_arcs.cuh: sweep, phase C, pass 1, pair settlement,
exact scan) is plain C++ that also compiles for the host.  tests/host/arcs_host.cpp builds it with g++ and this
test runs it serially on the CPU against the oracle's polar_targets (utils/tal.py:1257-1277) - the same source
the GPU runs, checked without a GPU: interior, exterior (in the box, outside the polygon) and near-contour
anchors, 36 and 72 rays, a circle known-answer case, a contour vertex on the anchor."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from util import ROOT, synth

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import polar_oracle as O  # noqa: E402


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("arcs") / "arcs_host.so")
    cmd = ["g++", "-O2", "-shared", "-fPIC", "-I/usr/local/cuda/include", "-o", out,
           os.path.join(ROOT, "tests", "host", "arcs_host.cpp")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lib = C.CDLL(out)
    lib.arcs_polar_targets.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.arcs_polar_targets.restype = C.c_int
    return lib


def run_host(lib, anchors, contours, R):
    anchors = np.ascontiguousarray(anchors, np.float32)
    contours = np.ascontiguousarray(contours, np.float32)
    M = anchors.shape[0]
    out = np.zeros((M, R), np.float32)
    stats = np.zeros(8, np.int64)
    assert lib.arcs_polar_targets(anchors.ctypes.data, contours.ctypes.data, M, R, out.ctypes.data, stats.ctypes.data) == 0
    return out, stats


def candidates(cfg, batch, limit):
    """in-box anchors of every GT with that GT's contour (px)"""
    anc, st = O.make_anchors(cfg.level_shapes, cfg.strides)
    anc_px = anc * st
    A, Cc, n, seen = [], [], 0, {}
    for bi, b in enumerate(batch["batch_idx"].long().tolist()):
        k = seen.get(b, 0)
        seen[b] = k + 1
        seg = batch["segments"][b][k] * cfg.imgsz
        (x0, y0), (x1, y1) = seg.min(0)[0], seg.max(0)[0]
        m = (anc_px[:, 0] > x0) & (anc_px[:, 0] < x1) & (anc_px[:, 1] > y0) & (anc_px[:, 1] < y1)
        a = anc_px[m]
        A.append(a)
        Cc.append(seg[None].expand(a.shape[0], -1, -1))
        n += a.shape[0]
        if n > limit:
            break
    return torch.cat(A).contiguous().float(), torch.cat(Cc).contiguous().float()


def check(lib, A, Cc, R):
    out, stats = run_host(lib, A.numpy(), Cc.numpy(), R)
    ref = O.polar_targets(A, Cc, R)
    t, amb = ref["t"].numpy(), ref["ambiguous"].numpy()
    rel = np.abs(out - t) / np.maximum(np.abs(t), 1e-6)
    assert int(((rel > 1e-5) & ~amb).sum()) == 0          # 1e-5 relative wherever the reference's pick is tie-free
    lo, hi = ref["t_lo"].numpy(), ref["t_hi"].numpy()
    assert bool(((out >= lo * (1 - 1e-5)) & (out <= hi * (1 + 1e-5))).all())   # ambiguous rays: inside the envelope
    assert amb.mean() < 0.01
    return stats


@pytest.mark.parametrize("seed,R,imgsz,gts", [(1, 36, 640, 20), (7, 36, 640, 12), (3, 72, 1280, 30), (11, 72, 640, 8)])
def test_host_arcs_match_oracle(host_lib, seed, R, imgsz, gts):
    cfg = synth.PathConfig("t", 2, gts, imgsz, rays=R)
    batch = synth.make_gts(cfg, seed)
    A, Cc = candidates(cfg, batch, 6000)
    stats = check(host_lib, A, Cc, R)
    cand = stats[0]
    assert cand == A.shape[0]
    # the exact 360-point scan must stay the exception, or the kernel's speed is gone
    assert stats[4] / cand < 0.5, f"exact scans per candidate: {stats[4] / cand:.3f}"
    assert stats[6] == 0


def test_host_arcs_circle_known_answer(host_lib):
    """circle of radius 50 centred on anchor (324,324) -> every ray target is 50."""
    t = np.linspace(0, 2 * np.pi, 360, endpoint=False)
    poly = np.stack([324 + 50 * np.cos(t), 324 + 50 * np.sin(t)], 1) / 640.0
    seg = synth.resample_closed(poly) * 640.0
    out, _ = run_host(host_lib, np.array([[324.0, 324.0]]), seg[None], 36)
    assert np.all(np.abs(out - 50.0) < 5e-3)   # resampling puts the points on chords of the 360-gon
    ref = O.polar_targets(torch.tensor([[324.0, 324.0]]), torch.from_numpy(seg)[None], 36)["t"].numpy()
    assert np.all(np.abs(out - ref) <= 1e-5 * ref)


def test_host_arcs_vertex_on_anchor_and_far_outside(host_lib):
    """A contour vertex exactly on the anchor (atan2(0,0) = 0 in the reference), and anchors far outside the
    polygon (every ray gated or crossed twice).  The polygon has no axis-parallel edge through an anchor, which
    would put a whole edge at the same angle (the reference's pick among tied points is unspecified)."""
    tri = np.array([[100.0, 100.0], [150.0, 117.3], [112.7, 160.0]]) / 640.0
    seg = torch.from_numpy(synth.resample_closed(tri) * 640.0)
    assert float(seg[0, 0]) == 100.0 and float(seg[0, 1]) == 100.0
    anchors = torch.tensor([[100.0, 100.0], [124.0, 124.0], [60.0, 60.0], [300.0, 120.0], [116.0, 108.0], [140.0, 132.0]])
    Cc = seg[None].expand(anchors.shape[0], -1, -1).contiguous()
    ref = O.polar_targets(anchors, Cc, 36)
    out, stats = run_host(host_lib, anchors.numpy(), Cc.numpy(), 36)
    t, amb = ref["t"].numpy(), ref["ambiguous"].numpy()
    rel = np.abs(out - t) / np.maximum(np.abs(t), 1e-6)
    assert int(((rel > 1e-5) & ~amb).sum()) == 0
    assert amb.sum() < 8
