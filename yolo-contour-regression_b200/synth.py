"""Seeded synthetic inputs for the polar-contour hot path.

Everything here is numpy (PCG64, stable across machines) + torch CPU tensors; nothing is read from
the reference at run time.  The contour resampler restates the wire format produced by the
reference's data pipeline (`ultralytics/utils/ops.py:676-693`: close the polygon, then
`np.interp` onto `linspace(0, len, n)`), so point 0 == point n-1 exactly as in real batches.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch

CONTOUR_POINTS = 360  # fixed by the reference wire format (utils/instance.py:202)


@dataclass
class PathConfig:
    """One workload of BASELINE.json (`configs`)."""
    name: str
    batch: int
    gts: int
    imgsz: int
    rays: int = 36
    nc: int = 80
    strides: tuple = (8, 16, 32)

    @property
    def level_shapes(self):
        return [(self.imgsz // s, self.imgsz // s) for s in self.strides]

    @property
    def anchors(self):
        return sum(h * w for h, w in self.level_shapes)


CONFIGS = {
    "C1": PathConfig("C1", 2, 8, 640),
    "C2": PathConfig("C2", 64, 20, 640),
    "C3": PathConfig("C3", 256, 0, 640),
    "C4": PathConfig("C4", 32, 200, 1280, rays=72),
    "C5": PathConfig("C5", 512, 20, 640),
}


def resample_closed(poly: np.ndarray, n: int = CONTOUR_POINTS) -> np.ndarray:
    """(m,2) open polygon -> (n,2) float32, closed then linearly resampled."""
    s = np.concatenate((poly, poly[0:1, :]), axis=0)
    x = np.linspace(0, len(s) - 1, n)
    xp = np.arange(len(s))
    return np.stack([np.interp(x, xp, s[:, k]) for k in range(2)], axis=1).astype(np.float32)


def make_gts(cfg: PathConfig, seed: int, ragged: bool = False, gts: int | None = None):
    """Star-shaped polygons, normalised coordinates.  Returns the reference `batch` dict fields.

    ragged=True draws a different GT count per image (including zero for image 1 when B>1) so the
    padding / mask_gt logic is exercised."""
    rng = np.random.default_rng(seed)
    G = cfg.gts if gts is None else gts
    t = np.linspace(0.0, 2.0 * math.pi, CONTOUR_POINTS, endpoint=False)
    batch_idx, cls, boxes, segs = [], [], [], []
    for b in range(cfg.batch):
        n = G
        if ragged:
            n = int(rng.integers(1, G + 1))
            if b == 1:
                n = 0
        img_segs = []
        for _ in range(n):
            # centre with a sub-pixel irrational jitter so no contour point sits on an anchor centre
            c = rng.uniform(0.15, 0.85, size=2) + np.array([math.sqrt(2.0), math.sqrt(3.0)]) * 1e-4
            r0 = rng.uniform(0.03, 0.15)
            phi = rng.uniform(0, 2 * math.pi)
            lobes = int(rng.integers(2, 6))
            r = r0 * (1.0 + 0.25 * np.sin(lobes * t + phi))
            poly = np.stack([c[0] + r * np.cos(t), c[1] + r * np.sin(t)], axis=1)
            poly = np.clip(poly, 0.0, 1.0)
            seg = resample_closed(poly)
            x0, y0 = seg.min(0)
            x1, y1 = seg.max(0)
            boxes.append([(x0 + x1) / 2, (y0 + y1) / 2, x1 - x0, y1 - y0])
            cls.append(float(rng.integers(0, cfg.nc)))
            batch_idx.append(float(b))
            img_segs.append(seg)
        segs.append(torch.from_numpy(np.stack(img_segs)) if img_segs
                    else torch.zeros(0, CONTOUR_POINTS, 2))
    return {
        "batch_idx": torch.tensor(batch_idx, dtype=torch.float32),
        "cls": torch.tensor(cls, dtype=torch.float32).view(-1, 1),
        "bboxes": torch.tensor(np.array(boxes, dtype=np.float32).reshape(-1, 4)),
        "segments": segs,
    }


def make_feats(cfg: PathConfig, seed: int, batch: int | None = None, hot_cells: int = 20):
    """Head outputs per level, `(B, R+nc, H_l, W_l)` fp32: first R channels rays (stride units),
    then nc class logits.  Logits are strongly negative with +7 bumps on 3x3 neighbourhoods of
    `hot_cells` random cells per image and level-0, so a few hundred anchors pass conf 0.25."""
    rng = np.random.default_rng(seed + 7919)
    B = cfg.batch if batch is None else batch
    feats = []
    for li, (h, w) in enumerate(cfg.level_shapes):
        rays = np.abs(rng.standard_normal((B, cfg.rays, h, w), dtype=np.float32)) * 2.0 + 2.0
        mean = rng.uniform(-6.0, -3.0, size=(B, cfg.nc, 1, 1)).astype(np.float32)
        logit = rng.standard_normal((B, cfg.nc, h, w), dtype=np.float32) + mean
        n_hot = max(1, hot_cells // (2 ** li))
        for b in range(B if (h > 2 and w > 2) else 0):   # (grids of 2x2 cells have no interior cell)
            ys = rng.integers(1, h - 1, size=n_hot)
            xs = rng.integers(1, w - 1, size=n_hot)
            cs = rng.integers(0, cfg.nc, size=n_hot)
            for y, x, c in zip(ys, xs, cs):
                logit[b, c, y - 1:y + 2, x - 1:x + 2] += 7.0
        feats.append(torch.from_numpy(np.concatenate([rays, logit], axis=1)))
    return feats


def make_feats_near_gt(cfg: PathConfig, seed: int, batch_dict) -> list:
    """Like make_feats, but anchors near GT centres predict rays close to the true contour radius and
    a raised logit at the GT class, so assignment looks like mid-training rather than step 0."""
    feats = make_feats(cfg, seed, hot_cells=0)
    rng = np.random.default_rng(seed + 104729)
    bi = batch_dict["batch_idx"].long().tolist()
    for n, b in enumerate(bi):
        cx, cy, bw, bh = (batch_dict["bboxes"][n] * cfg.imgsz).tolist()
        c = int(batch_dict["cls"][n].item())
        rad = 0.25 * (bw + bh)
        for li, s in enumerate(cfg.strides):
            h, w = cfg.level_shapes[li]
            x0, x1 = max(0, int((cx - bw / 4) / s)), min(w, int((cx + bw / 4) / s) + 1)
            y0, y1 = max(0, int((cy - bh / 4) / s)), min(h, int((cy + bh / 4) / s) + 1)
            if x1 <= x0 or y1 <= y0:
                continue
            shape = (cfg.rays, y1 - y0, x1 - x0)
            noise = rng.standard_normal(shape).astype(np.float32) * 0.15 + 1.0
            feats[li][b, :cfg.rays, y0:y1, x0:x1] = torch.from_numpy(noise * (rad / s))
            bump = rng.uniform(2.0, 6.0, size=shape[1:]).astype(np.float32)
            feats[li][b, cfg.rays + c, y0:y1, x0:x1] += torch.from_numpy(bump)
    return feats
