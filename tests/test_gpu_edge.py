"""GPU: degenerate inputs the reference's data pipeline can produce and the synthetic generators never do.

* a contour point exactly on an anchor centre: `atan2(0, 0) = 0`, so the reference files that point under ray 0 at
  distance 0 (utils/tal.py:1286-1301);
* an instance whose polygon was clipped away - a real box with an all-zero contour (360 identical points at the
  origin: every angle ties, one angular bin takes all 360 points);
* two identical instances in one image: every anchor both pick has bit-identical overlaps, `argmax` keeps the first
  (utils/tal.py:214-248).
"""
import numpy as np
import pytest
import torch

from util import rel_err, synth
from oracle import polar_oracle as po

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _inputs(cfg, feats, batch, dev):
    B = feats[0].shape[0]
    no = cfg.rays + cfg.nc
    cat = torch.cat([f.view(B, no, -1) for f in feats], 2)
    rays, logits = cat.split((cfg.rays, cfg.nc), 1)
    logits = logits.permute(0, 2, 1).contiguous()
    rays = rays.permute(0, 2, 1).contiguous()
    shapes = [tuple(f.shape[2:]) for f in feats]
    anc, st = po.make_anchors(shapes, cfg.strides)
    targets = po.pack_targets(batch, B, (cfg.imgsz, cfg.imgsz))
    gl, gb, gc = targets.split((1, 4, 720), 2)
    mask_gt = (gb.sum(2, keepdim=True) > 0).float()
    cpu = dict(scores=logits.sigmoid(), rays=rays * st, anc=anc * st, gl=gl, gb=gb, mask_gt=mask_gt, gc=gc, st=st)
    return cpu, {k: v.to(dev) for k, v in cpu.items()}, shapes


def _run(cfg, cpu, gpu, shapes):
    from ycr_b200.tal import TaskAlignedAssigner
    asg = TaskAlignedAssigner(topk=10, num_classes=cfg.nc, alpha=0.5, beta=4.0)
    asg.debug_metrics = True
    out = asg(gpu["scores"], gpu["rays"], gpu["anc"], gpu["gl"], gpu["gb"], gpu["mask_gt"], gpu["gc"], gpu["st"],
              None, 0, None, grid=(shapes, list(cfg.strides)))
    ref = po.assign(cpu["scores"], cpu["rays"], cpu["anc"], cpu["gl"], cpu["gb"], cpu["mask_gt"], cpu["gc"])
    return asg, [t.cpu() for t in out], ref


def _rebox(batch, n):
    seg = torch.cat(batch["segments"])[n]
    x0, y0 = seg.min(0)[0]
    x1, y1 = seg.max(0)[0]
    batch["bboxes"][n] = torch.tensor([(x0 + x1) / 2, (y0 + y1) / 2, x1 - x0, y1 - y0])


def test_contour_points_on_anchor_centres():
    dev = torch.device("cuda:0")
    cfg = synth.PathConfig("on", 1, 2, 128, nc=4)      # 128 px: k/128 is exact, anchors at 4+8k, 8+16k, 16+32k
    batch = synth.make_gts(cfg, 77)
    segs = batch["segments"][0]
    hit = []
    for n in range(2):
        px = segs[n] * 128.0
        for j, (stride, off) in zip((5, 130, 250), ((8, 4), (16, 8), (32, 16))):
            snapped = torch.round((px[j] - off) / stride) * stride + off   # the anchor centre nearest to contour point j
            segs[n, j] = snapped / 128.0
            hit.append((n, stride, snapped.tolist()))
        _rebox(batch, n)
    feats = synth.make_feats_near_gt(cfg, 77, batch)
    cpu, gpu, shapes = _inputs(cfg, feats, batch, dev)
    # the snapped points really coincide with anchors, bit for bit, after the packing arithmetic
    anc = cpu["anc"]
    coin = 0
    for n in range(2):
        c = cpu["gc"][0, n].view(360, 2)
        coin += int((c[:, None, :] == anc[None, :, :]).all(2).any(1).sum())
    assert coin >= 6
    asg, out, ref = _run(cfg, cpu, gpu, shapes)
    tl, tb, ts, mp, tgi, gd, cen, fg = out
    ov = asg.last_overlaps.cpu()
    lo, hi = ref["overlaps_lo"], ref["overlaps_hi"]
    assert torch.equal(ov != 0, ref["overlaps"] != 0)
    assert bool(((ov >= lo * (1 - TOL)) & (ov <= hi * (1 + TOL))).all())
    sure = lo == hi
    assert float(sure.float().mean()) > 0.95
    assert rel_err(ov[sure], ref["overlaps"][sure]) < TOL
    # the candidates that sit ON a contour point are among the compared ones
    on = torch.zeros_like(sure)
    for n in range(2):
        c = cpu["gc"][0, n].view(360, 2)
        on[0, n] = (c[:, None, :] == anc[None, :, :]).all(2).any(0)
    assert int((on & sure & (ref["overlaps"] != 0)).sum()) >= 1
    if bool(ref["certain"].all()):
        assert torch.equal(tgi, ref["target_gt_idx"]) and torch.equal(fg, ref["fg_mask"]) and torch.equal(mp, ref["mask_pos"])
        ok = ~ref["gt_dist_ambiguous"]
        assert rel_err(gd[ok], ref["gt_dist"][ok]) < TOL


def test_real_box_with_all_zero_contour():
    dev = torch.device("cuda:0")
    cfg = synth.PathConfig("z", 2, 3, 160, nc=5)
    batch = synth.make_gts(cfg, 91)
    batch["segments"][0][1].zero_()          # image 0, GT 1: polygon clipped away, box kept
    batch["segments"][1][0].zero_()
    feats = synth.make_feats_near_gt(cfg, 91, batch)
    cpu, gpu, shapes = _inputs(cfg, feats, batch, dev)
    assert float(cpu["gc"][0, 1].abs().max()) == 0.0 and float(cpu["gb"][0, 1].sum()) > 0
    asg, out, ref = _run(cfg, cpu, gpu, shapes)
    tl, tb, ts, mp, tgi, gd, cen, fg = out
    ov = asg.last_overlaps.cpu()
    lo, hi = ref["overlaps_lo"], ref["overlaps_hi"]
    assert torch.equal(ov != 0, ref["overlaps"] != 0)
    assert bool(((ov >= lo * (1 - TOL)) & (ov <= hi * (1 + TOL))).all())
    sure = lo == hi
    z = torch.zeros_like(sure)
    z[0, 1] = True
    z[1, 0] = True
    assert int((z & sure & (ref["overlaps"] != 0)).sum()) >= 10     # candidates of the degenerate GTs are compared
    assert rel_err(ov[sure], ref["overlaps"][sure]) < TOL
    # whatever the all-ties do to the oracle's certificate, finite outputs and the loss path must hold
    from ycr_b200.loss import v8SegmentationLoss
    crit = v8SegmentationLoss(nc=cfg.nc, nm=cfg.rays, strides=cfg.strides, device=dev)
    fgd = [f.to(dev).requires_grad_(True) for f in feats]
    total, items = crit((fgd, 5, 2), batch)
    total.backward()
    assert bool(torch.isfinite(total)) and all(bool(torch.isfinite(f.grad).all()) for f in fgd)
    if bool(ref["certain"].all()):
        assert torch.equal(tgi, ref["target_gt_idx"]) and torch.equal(mp, ref["mask_pos"])
        r = po.seg_loss(feats, batch, cfg.strides, cfg.nc, cfg.rays)
        assert rel_err(items.cpu(), r["loss_items"]) < TOL


def test_identical_instances_resolve_to_the_first():
    dev = torch.device("cuda:0")
    cfg = synth.PathConfig("dup", 1, 3, 160, nc=6)
    batch = synth.make_gts(cfg, 55)
    batch["segments"][0][2] = batch["segments"][0][0].clone()    # GT 2 := GT 0 (same contour, box and class)
    batch["bboxes"][2] = batch["bboxes"][0].clone()
    batch["cls"][2] = batch["cls"][0].clone()
    feats = synth.make_feats_near_gt(cfg, 55, batch)
    cpu, gpu, shapes = _inputs(cfg, feats, batch, dev)
    asg, out, ref = _run(cfg, cpu, gpu, shapes)
    tl, tb, ts, mp, tgi, gd, cen, fg = out
    ov = asg.last_overlaps.cpu()
    assert torch.equal(ov[0, 0], ov[0, 2])                       # bit-identical metrics for the twins
    # both twins pick the same ten anchors; all of them go to the first (argmax keeps the first maximum)
    both = mp[0, 0].bool() | mp[0, 2].bool()
    assert int(mp[0, 2].sum()) == 0 and int(mp[0, 0].sum()) >= 1
    assert not bool((tgi[0][both] == 2).any())
    assert torch.equal(tgi, ref["target_gt_idx"])
    assert torch.equal(fg, ref["fg_mask"])
    assert torch.equal(mp, ref["mask_pos"])
    ok = ~ref["gt_dist_ambiguous"]
    assert rel_err(gd[ok], ref["gt_dist"][ok]) < TOL
    nz = ref["target_scores"] != 0
    assert torch.equal(ts != 0, nz)
    assert rel_err(ts[nz], ref["target_scores"][nz]) < TOL
