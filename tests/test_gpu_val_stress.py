"""GPU: the validator boundary (postprocess / update_metrics) and the stress-shaped configuration
(1280 px, many GTs, 72 rays — BASELINE.json configs[3] at reduced batch) against the oracle."""
import numpy as np
import pytest
import torch

from util import rel_err
from oracle import polar_oracle as po

pytestmark = pytest.mark.gpu


def _dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def test_validator_postprocess_and_metrics():
    """SegmentationValidator.postprocess = NMS(multi_label=True) (models/yolo/segment/val.py:46-61);
    update_metrics appends (correct_bboxes, correct_masks, conf, cls, target cls) per image (:149-219)."""
    from ycr_b200.val import SegmentationValidator
    from ycr_b200.head import decode
    from ycr_b200 import synth
    dev = _dev()
    cfg = synth.PathConfig("v", 3, 5, 320, nc=10)
    batch = synth.make_gts(cfg, 71)
    feats = synth.make_feats_near_gt(cfg, 71, batch)
    allpred = decode([f.to(dev) for f in feats], cfg.strides, cfg.nc, cfg.rays)
    val = SegmentationValidator(nc=cfg.nc, device=dev)
    val.args.conf = 0.25
    preds = val.postprocess((allpred, None))
    ref, margin = po.nms(po.decode(feats, cfg.strides, cfg.nc, cfg.rays), 0.25, 0.7, multi_label=True, nc=cfg.nc,
                         max_det=300)
    assert margin > 1e-5
    assert [p.shape[0] for p in preds] == [r.shape[0] for r in ref]
    for p, r in zip(preds, ref):
        assert torch.equal(p[:, 5].cpu(), r[:, 5])
        assert bool(((p.cpu() - r).abs() <= 1e-5 * r.abs() + 1e-5).all())
    b = dict(batch)
    b["img"] = torch.zeros(cfg.batch, 3, cfg.imgsz, cfg.imgsz)
    val.update_metrics(preds, b)
    assert val.seen == cfg.batch and len(val.stats) == cfg.batch
    for (cb, cm, conf, pcls, tcls), p in zip(val.stats, preds):
        assert cb.shape == (p.shape[0], 10) and cm.shape == (p.shape[0], 10)
        assert not bool(cm.any())                      # masks are all-zero in the reference snapshot
        assert torch.equal(conf, p[:, 4]) and tcls.numel() == cfg.gts
        # at most one detection is matched to a GT per threshold (utils matching is one-to-one)
        assert int(cb[:, 0].sum()) <= cfg.gts and int(cb[:, -1].sum()) <= int(cb[:, 0].sum())
    # every image has at least one near-GT anchor predicted with the right class in this synthetic set-up
    assert sum(int(s[0][:, 0].sum()) for s in val.stats) > 0


def test_stress_shape_rays72_1280():
    """configs[3] shape at reduced batch: 1280 px (A = 33600), many GTs per image, 72 rays."""
    from ycr_b200.loss import v8SegmentationLoss
    from ycr_b200 import synth
    dev = _dev()
    cfg = synth.PathConfig("c4s", 1, 24, 1280, rays=72, nc=80)
    for seed in (81, 82, 83, 84):
        batch = synth.make_gts(cfg, seed)
        # keep the oracle affordable: shrink the objects (candidates scale with box area)
        batch["bboxes"][:, 2:] *= 0.45
        segs = batch["segments"][0]
        c = batch["bboxes"][:, None, :2]
        batch["segments"] = [(segs - c) * 0.45 + c]
        feats = synth.make_feats_near_gt(cfg, seed, batch)
        ref = po.seg_loss(feats, batch, cfg.strides, cfg.nc, cfg.rays)
        if bool(ref["assign"]["certain"].all()):
            break
    else:
        pytest.skip("no tie-free seed")
    crit = v8SegmentationLoss(nc=cfg.nc, nm=72, strides=cfg.strides, device=dev)
    fg = [f.to(dev).requires_grad_(True) for f in feats]
    total, items = crit((fg, 5, 2), batch)
    total.backward()
    assert rel_err(items.cpu(), ref["loss_items"]) < 1e-5
    for f, r in zip(fg, ref["grads"]):
        assert float((f.grad.cpu() - r).abs().max()) <= 1e-5 * float(r.abs().max())


def test_many_gts_per_image_200():
    """G = 200 padded GTs per image (the stress config's GT count) on a 640 image, 36 rays: exercises the
    per-image resolution kernel with thousands of positives and heavy multi-GT overlap."""
    from ycr_b200.tal import TaskAlignedAssigner
    from ycr_b200 import synth
    dev = _dev()
    cfg = synth.PathConfig("g200", 1, 200, 640, nc=20)
    batch = synth.make_gts(cfg, 91)
    batch["bboxes"][:, 2:] *= 0.3
    segs = batch["segments"][0]
    c = batch["bboxes"][:, None, :2]
    batch["segments"] = [(segs - c) * 0.3 + c]
    feats = synth.make_feats_near_gt(cfg, 91, batch)
    B, no = 1, cfg.rays + cfg.nc
    cat = torch.cat([f.view(B, no, -1) for f in feats], 2)
    rays, logits = cat.split((cfg.rays, cfg.nc), 1)
    anc, st = po.make_anchors(cfg.level_shapes, cfg.strides)
    t = po.pack_targets(batch, B, (640, 640))
    gl, gb, gc = t.split((1, 4, 720), 2)
    mg = (gb.sum(2, keepdim=True) > 0).float()
    scores, prays = logits.permute(0, 2, 1).contiguous().sigmoid(), rays.permute(0, 2, 1).contiguous() * st
    ref = po.assign(scores, prays, anc * st, gl, gb, mg, gc)
    asg = TaskAlignedAssigner(topk=10, num_classes=cfg.nc, alpha=0.5, beta=4.0)
    out = asg(scores.to(dev), prays.to(dev), (anc * st).to(dev), gl.to(dev), gb.to(dev), mg.to(dev), gc.to(dev),
              st.to(dev), None, 0, None, grid=(cfg.level_shapes, list(cfg.strides)))
    tl, tb, ts, mp, tgi, gd, cen, fgm = [x.cpu() for x in out]
    if not bool(ref["certain"].all()):
        pytest.skip("seed not tie-free")
    assert torch.equal(tgi, ref["target_gt_idx"]) and torch.equal(fgm, ref["fg_mask"]) and torch.equal(mp, ref["mask_pos"])
    ok = ~ref["gt_dist_ambiguous"]
    assert rel_err(gd[ok], ref["gt_dist"][ok]) < 1e-5
    nz = ref["target_scores"] != 0
    assert rel_err(ts[nz], ref["target_scores"][nz]) < 1e-5
