"""How often is the reference's own ray target ambiguous?  The reference picks the four contour points nearest in
angle to each ray from fp32 atan2 degrees; where the 4th/5th gap or the distance of the nearest point to the
3-degree gate is below 2e-4 degrees (a few fp32 ulps at 360 degrees) its result depends on rounding, and the parity
tests do not pin it.  This script measures that fraction with the oracle on the synthetic C2 / C4 data (CPU only):
    python tools/ambiguity_report.py > profiles/r02_ambiguity_band.json"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ycr_b200  # noqa: E402,F401
from ycr_b200 import synth  # noqa: E402
from oracle import polar_oracle as po  # noqa: E402


def measure(name, images, gts_per_image, seed):
    cfg = synth.CONFIGS[name]
    sub = synth.PathConfig(name, images, cfg.gts, cfg.imgsz, rays=cfg.rays, nc=cfg.nc)
    batch = synth.make_gts(sub, seed)
    anc, st = po.make_anchors(sub.level_shapes, sub.strides)
    anc = anc * st
    n_rays = n_amb = n_cand = 0
    for b in range(images):
        segs = batch["segments"][b][:gts_per_image] * cfg.imgsz
        for seg in segs:
            (x0, y0), (x1, y1) = seg.min(0)[0], seg.max(0)[0]
            m = (anc[:, 0] > x0) & (anc[:, 0] < x1) & (anc[:, 1] > y0) & (anc[:, 1] < y1)
            a = anc[m]
            if a.shape[0] == 0:
                continue
            pt = po.polar_targets(a, seg[None].expand(a.shape[0], -1, -1).contiguous(), cfg.rays)
            n_cand += a.shape[0]
            n_rays += pt["ambiguous"].numel()
            n_amb += int(pt["ambiguous"].sum())
    return {"config": name, "rays": cfg.rays, "imgsz": cfg.imgsz, "candidates": n_cand, "candidate_rays": n_rays,
            "ambiguous_rays": n_amb, "fraction": n_amb / max(n_rays, 1), "band_deg": 2e-4}


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count() or 1)
    out = [measure("C2", 3, 20, 301), measure("C4", 1, 12, 302)]
    print(json.dumps(out, indent=1))
