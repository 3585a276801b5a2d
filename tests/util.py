"""Shared helpers for the parity tests: rebuild the seeded inputs a golden file was made from."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import ycr_b200  # noqa: E402,F401
from ycr_b200 import synth  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))


def cfg_of(g):
    b, gts, imgsz, rays, nc = [int(v) for v in g["cfg"]]
    return synth.PathConfig("golden", b, gts, imgsz, rays=rays, nc=nc)


def train_inputs(g):
    """(cfg, feats, batch) exactly as tests/golden/make_golden.py built them."""
    cfg = cfg_of(g)
    seed = int(g["seed"])
    batch = synth.make_gts(cfg, seed, ragged=bool(g["ragged"]))
    feats = synth.make_feats_near_gt(cfg, seed, batch) if bool(g["near"]) else synth.make_feats(cfg, seed)
    if "feat0" in g:  # the stored copy must be what the generator reproduces
        for li, f in enumerate(feats):
            assert np.array_equal(f.numpy(), g[f"feat{li}"]), "synthetic generator drifted"
        assert np.array_equal(torch.cat(batch["segments"]).numpy(), g["segments"])
    else:
        dig = np.array([float(f.double().sum()) for f in feats])
        assert np.allclose(dig, g["feat_digest"], rtol=0, atol=0), "synthetic generator drifted"
    return cfg, feats, batch


def infer_inputs(g):
    cfg = cfg_of(g)
    feats = synth.make_feats(cfg, int(g["seed"]))
    if "feat0" in g:
        for li, f in enumerate(feats):
            assert np.array_equal(f.numpy(), g[f"feat{li}"])
    return cfg, feats


def rel_err(a, b, floor=1e-12):
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    return float(((a - b).abs() / b.abs().clamp(min=floor)).max()) if a.numel() else 0.0


def split_rows(rows, counts):
    out, s = [], 0
    for c in counts:
        out.append(rows[s:s + int(c)])
        s += int(c)
    return out
