"""GPU parity tests of the training path (assigner, polar targets, fused loss fwd+bwd), all through
the C-ABI library via the host mirror.  Reference = golden vectors produced by the reference itself
(tests/golden) and the oracle on the same seeded inputs.

Bars (BASELINE.json north_star): bit-exact for assigned GT indices, foreground masks, top-k selections
(on tie-free inputs, which the oracle's margin checker certifies); 1e-5 relative in fp32 for polar
targets, loss terms and gradients."""
import numpy as np
import pytest
import torch

from util import load_golden, train_inputs, rel_err
from oracle import polar_oracle as po

pytestmark = pytest.mark.gpu

TOL = 1e-5


def _dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _assigner_inputs(cfg, feats, batch, dev):
    """What utils/loss.py:815-862 feeds the assigner, computed with plain torch on the device."""
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
    ss = [torch.full((h * w, 1), float(s)) for (h, w), s in zip(shapes, cfg.strides)]
    cpu = dict(scores=logits.sigmoid(), rays=rays * st, anc=anc * st, gl=gl, gb=gb, mask_gt=mask_gt, gc=gc, st=st)
    gpu = {k: v.to(dev) for k, v in cpu.items()}
    gpu["ss"] = [s.to(dev) for s in ss]
    return cpu, gpu, shapes


def _check_assign(out, ref, nc):
    """out: product 8-tuple (+ dense metrics); ref: oracle dict.  Only images the oracle certifies."""
    tl, tb, ts, mp, tgi, gd, cen, fg = [t.cpu() for t in out]
    cert = ref["certain"]
    assert bool(cert.all()), "test inputs must be tie-free"
    assert torch.equal(tgi, ref["target_gt_idx"])
    assert torch.equal(fg, ref["fg_mask"])
    assert torch.equal(mp, ref["mask_pos"])
    assert torch.equal(tl, ref["target_labels"])
    assert torch.equal(tb, ref["target_bboxes"])
    assert gd.shape == ref["gt_dist"].shape
    ok = ~ref["gt_dist_ambiguous"]
    assert float(ok.float().mean()) > 0.99 if ok.numel() else True
    assert rel_err(gd[ok], ref["gt_dist"][ok]) < TOL
    clean = ok.all(1)
    assert rel_err(cen[clean], ref["centerness"][clean]) < TOL
    nz = ref["target_scores"] != 0
    assert torch.equal(ts != 0, nz)
    assert rel_err(ts[nz], ref["target_scores"][nz]) < 1e-5


@pytest.mark.parametrize("name", ["train_s160", "train_s320_ragged", "train_c1", "train_c1_near"])
def test_loss_matches_reference_golden(name):
    from ycr_b200.loss import v8SegmentationLoss
    dev = _dev()
    g = load_golden(name)
    cfg, feats, batch = train_inputs(g)
    crit = v8SegmentationLoss(nc=cfg.nc, nm=cfg.rays, strides=cfg.strides, device=dev)
    fg = [f.to(dev).requires_grad_(True) for f in feats]
    total, items = crit((fg, 5, 2), batch)
    total.backward()
    assert items.shape == (2,) and not items.requires_grad
    assert rel_err(total.detach().cpu(), g["loss"]) < TOL
    assert rel_err(items.cpu(), g["loss_items"]) < TOL
    for li, f in enumerate(fg):
        gr = f.grad.cpu()
        ref_s = torch.from_numpy(g[f"grad{li}_sample"])
        scale = float(ref_s.abs().max())
        assert float((gr.flatten()[::97] - ref_s).abs().max()) <= TOL * scale
        assert abs(float(gr.double().abs().sum()) - float(g[f"grad{li}_abssum"])) <= TOL * float(g[f"grad{li}_abssum"])
        nzi = torch.nonzero(gr[:, :36].flatten()).flatten().numpy()
        assert np.array_equal(nzi, g[f"grad{li}_ray_nz_idx"])          # exactly the positives get ray grads
        nzv = gr[:, :36].flatten()[nzi]
        assert float((nzv - torch.from_numpy(g[f"grad{li}_ray_nz_val"])).abs().max()) <= TOL * scale
        if f"grad{li}" in g:
            ref = torch.from_numpy(g[f"grad{li}"])
            assert float((gr - ref).abs().max()) <= TOL * float(ref.abs().max())


@pytest.mark.parametrize("name", ["train_s160", "train_s320_ragged", "train_c1", "train_c1_near"])
def test_assigner_matches_reference_golden(name):
    from ycr_b200.tal import TaskAlignedAssigner
    dev = _dev()
    g = load_golden(name)
    cfg, feats, batch = train_inputs(g)
    cpu, gpu, shapes = _assigner_inputs(cfg, feats, batch, dev)
    asg = TaskAlignedAssigner(topk=10, num_classes=cfg.nc, alpha=0.5, beta=4.0)
    asg.debug_metrics = True
    out = asg(gpu["scores"], gpu["rays"], gpu["anc"], gpu["gl"], gpu["gb"], gpu["mask_gt"], gpu["gc"], gpu["st"],
              gpu["ss"], 0, torch.tensor([float(cfg.imgsz)] * 2, device=dev))
    ref = po.assign(cpu["scores"], cpu["rays"], cpu["anc"], cpu["gl"], cpu["gb"], cpu["mask_gt"], cpu["gc"])
    _check_assign(out, ref, cfg.nc)
    # golden (reference) fields directly
    tl, tb, ts, mp, tgi, gd, cen, fgm = [t.cpu() for t in out]
    assert np.array_equal(tgi.numpy(), g["asg_target_gt_idx"])
    assert np.array_equal(fgm.numpy(), g["asg_fg_mask"])
    assert np.array_equal(torch.nonzero(mp).numpy(), g["asg_mask_pos_nz"])
    assert rel_err(gd, g["asg_gt_dist"]) < TOL
    assert rel_err(cen, g["asg_centerness"]) < TOL
    assert rel_err(ts[ts != 0], g["asg_target_scores_nz_val"]) < 1e-5
    # dense Polar-IoU of every candidate (get_box_metrics_polar) inside the oracle's envelope
    ov = asg.last_overlaps.cpu()
    lo, hi = ref["overlaps_lo"], ref["overlaps_hi"]
    assert bool(((ov >= lo * (1 - TOL)) & (ov <= hi * (1 + TOL))).all())
    assert torch.equal(ov != 0, ref["overlaps"] != 0)          # exactly the in-box candidates
    sure = lo == hi
    assert rel_err(ov[sure], ref["overlaps"][sure]) < TOL
    al = asg.last_align_metric.cpu()
    assert rel_err(al[sure], ref["align_metric"][sure], floor=1e-30) < 5e-5  # ov**4


def test_kat_circle():
    from ycr_b200.tal import TaskAlignedAssigner
    dev = _dev()
    g = load_golden("kat_circle")
    shapes = [(80, 80), (40, 40), (20, 20)]
    anc, st = po.make_anchors(shapes, [8, 16, 32])
    asg = TaskAlignedAssigner(topk=10, num_classes=3, alpha=0.5, beta=4.0)
    out = asg(torch.from_numpy(g["pd_scores"]).to(dev), torch.from_numpy(g["pd_rays"]).to(dev), (anc * st).to(dev),
              torch.tensor([[[1.]]], device=dev), torch.tensor([[[274., 274., 374., 374.]]], device=dev),
              torch.ones(1, 1, 1, device=dev), torch.from_numpy(g["gt_coor"]).to(dev), st.to(dev), None, 0, None,
              grid=(shapes, [8, 16, 32]))
    tl, tb, ts, mp, tgi, gd, cen, fg = [t.cpu() for t in out]
    assert int(fg.sum()) == 10
    assert np.array_equal(fg.numpy(), g["asg_fg_mask"])
    assert rel_err(gd, g["asg_gt_dist"]) < TOL
    assert rel_err(ts, g["asg_target_scores"], floor=1e-6) < 1e-5
    ci = int(g["centre_anchor"])
    row = torch.nonzero(mp[0, 0]).flatten().tolist().index(ci)
    assert float(gd[row].min()) > 49.99 and float(gd[row].max()) < 50.01
    assert abs(float(cen[row]) - 1.0) < 1e-3


def test_empty_batch_and_empty_image():
    from ycr_b200.loss import v8SegmentationLoss
    from ycr_b200 import synth
    dev = _dev()
    cfg = synth.PathConfig("e", 2, 3, 160, nc=10)
    feats = synth.make_feats(cfg, 5)
    empty = {"batch_idx": torch.zeros(0), "cls": torch.zeros(0, 1), "bboxes": torch.zeros(0, 4),
             "segments": [torch.zeros(0, 360, 2)] * 2}
    crit = v8SegmentationLoss(nc=cfg.nc, nm=cfg.rays, strides=cfg.strides, device=dev)
    fg = [f.to(dev).requires_grad_(True) for f in feats]
    total, items = crit((fg, 5, 2), empty)
    total.backward()
    ref = po.seg_loss(feats, empty, cfg.strides, cfg.nc, cfg.rays)
    assert float(items[0]) == 0.0
    assert rel_err(items.cpu(), ref["loss_items"], floor=1e-6) < TOL
    for f, r in zip(fg, ref["grads"]):
        assert float((f.grad.cpu() - r).abs().max()) <= TOL * float(r.abs().max())
    # one image without GTs inside a non-empty batch
    batch = synth.make_gts(cfg, 9, ragged=True)
    assert int((batch["batch_idx"] == 1).sum()) == 0
    fg = [f.to(dev).requires_grad_(True) for f in feats]
    total, items = crit((fg, 5, 2), batch)
    ref = po.seg_loss(feats, batch, cfg.strides, cfg.nc, cfg.rays)
    assert bool(ref["assign"]["certain"].all())
    assert rel_err(items.cpu(), ref["loss_items"]) < TOL


def test_rays72_against_oracle():
    """72-ray variant (config C4's ray count) on a small image; the reference hard-codes 36, the oracle is
    proven identical to it at 36 and parameterised."""
    from ycr_b200.loss import v8SegmentationLoss
    from ycr_b200 import synth
    dev = _dev()
    cfg = synth.PathConfig("r72", 2, 5, 320, rays=72, nc=10)
    for seed in (31, 32, 33):
        batch = synth.make_gts(cfg, seed)
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
    assert rel_err(items.cpu(), ref["loss_items"]) < TOL
    for f, r in zip(fg, ref["grads"]):
        assert float((f.grad.cpu() - r).abs().max()) <= TOL * float(r.abs().max())


def test_loss_deterministic_and_grad_scale():
    from ycr_b200.loss import v8SegmentationLoss
    from ycr_b200 import synth
    dev = _dev()
    cfg = synth.PathConfig("d", 4, 6, 320, nc=20)
    batch = synth.make_gts(cfg, 41)
    feats = synth.make_feats_near_gt(cfg, 41, batch)
    crit = v8SegmentationLoss(nc=cfg.nc, nm=cfg.rays, strides=cfg.strides, device=dev)
    runs = []
    for scale in (1.0, 1.0, 128.0):
        fg = [f.to(dev).requires_grad_(True) for f in feats]
        total, items = crit((fg, 5, 2), batch)
        (total * scale).backward()
        runs.append((total.detach().clone(), [f.grad.clone() for f in fg]))
    assert torch.equal(runs[0][0], runs[1][0])
    for a, b in zip(runs[0][1], runs[1][1]):
        assert torch.equal(a, b)                                  # bit-identical run to run
    for a, c in zip(runs[0][1], runs[2][1]):
        assert torch.equal(a * 128.0, c)                          # upstream gradient honoured


def test_batch_scale_matches_oracle_on_sampled_images():
    """Config C2 shape (B=64, G=20, 640, nc=80): the product runs the full batch; the oracle re-derives
    the assignment of a few images (assignment is image-independent)."""
    from ycr_b200.tal import TaskAlignedAssigner
    from ycr_b200 import synth
    dev = _dev()
    cfg = synth.PathConfig("C2s", 16, 20, 640, nc=80)
    batch = synth.make_gts(cfg, 51)
    feats = synth.make_feats_near_gt(cfg, 51, batch)
    cpu, gpu, shapes = _assigner_inputs(cfg, feats, batch, dev)
    asg = TaskAlignedAssigner(topk=10, num_classes=cfg.nc, alpha=0.5, beta=4.0)
    out = asg(gpu["scores"], gpu["rays"], gpu["anc"], gpu["gl"], gpu["gb"], gpu["mask_gt"], gpu["gc"], gpu["st"],
              gpu["ss"], 0, None, grid=(shapes, list(cfg.strides)))
    tl, tb, ts, mp, tgi, gd, cen, fg = [t.cpu() for t in out]
    checked = 0
    for b in (0, 7, 15):
        sl = slice(b, b + 1)
        ref = po.assign(cpu["scores"][sl], cpu["rays"][sl], cpu["anc"], cpu["gl"][sl], cpu["gb"][sl],
                        cpu["mask_gt"][sl], cpu["gc"][sl])
        if not bool(ref["certain"][0]):
            continue
        checked += 1
        assert torch.equal(tgi[sl], ref["target_gt_idx"])
        assert torch.equal(fg[sl], ref["fg_mask"])
        assert torch.equal(mp[sl], ref["mask_pos"])
        nz = ref["target_scores"] != 0
        assert rel_err(ts[sl][nz], ref["target_scores"][nz]) < 1e-5
        r0 = int(mp[:b].sum())
        rows = gd[r0:r0 + int(mp[sl].sum())]
        ok = ~ref["gt_dist_ambiguous"]
        assert rel_err(rows[ok], ref["gt_dist"][ok]) < TOL
    assert checked >= 1


def test_bbox_loss_ciou_dfl_matches_reference():
    """Row 15 of SURVEY §8-a (dormant on the live path): BboxLoss forward + gradients vs the reference."""
    from ycr_b200.loss import BboxLoss
    dev = _dev()
    g = load_golden("bbox_loss")
    pd = torch.from_numpy(g["pred_dist"]).to(dev).requires_grad_(True)
    pb = torch.from_numpy(g["pred_bboxes"]).to(dev).requires_grad_(True)
    crit = BboxLoss(15, use_dfl=True)
    li, ld = crit(pd, pb, torch.from_numpy(g["anchor_points"]).to(dev), torch.from_numpy(g["target_bboxes"]).to(dev),
                  torch.from_numpy(g["target_scores"]).to(dev), float(g["tss"]), torch.from_numpy(g["fg_mask"]).to(dev))
    (li * float(g["gains"][0]) + ld * float(g["gains"][1])).backward()
    assert rel_err(li.detach().cpu(), g["loss_iou"]) < TOL
    assert rel_err(ld.detach().cpu(), g["loss_dfl"]) < TOL
    gd, gb = torch.from_numpy(g["grad_dist"]), torch.from_numpy(g["grad_bboxes"])
    assert float((pd.grad.cpu() - gd).abs().max()) <= TOL * float(gd.abs().max())
    assert float((pb.grad.cpu() - gb).abs().max()) <= TOL * float(gb.abs().max())
    assert torch.equal(pd.grad.cpu() != 0, gd != 0)                 # only foreground anchors get gradients
    # DFL switched off: loss_dfl is 0 and pred_dist gets no gradient
    pb2 = torch.from_numpy(g["pred_bboxes"]).to(dev).requires_grad_(True)
    li2, ld2 = BboxLoss(15, use_dfl=False)(pd.detach(), pb2, torch.from_numpy(g["anchor_points"]).to(dev),
                                           torch.from_numpy(g["target_bboxes"]).to(dev),
                                           torch.from_numpy(g["target_scores"]).to(dev), float(g["tss"]),
                                           torch.from_numpy(g["fg_mask"]).to(dev))
    assert float(ld2) == 0.0 and rel_err(li2.detach().cpu(), g["loss_iou"]) < TOL


def test_pack_targets_kernel_matches_oracle():
    """GT packing kernel (v8DetectionLoss.preprocess, utils/loss.py:215-239) incl. ragged and shuffled rows."""
    from ycr_b200.loss import v8SegmentationLoss
    from ycr_b200 import synth
    dev = _dev()
    cfg = synth.PathConfig("p", 5, 7, 320, nc=10)
    batch = synth.make_gts(cfg, 61, ragged=True)
    crit = v8SegmentationLoss(nc=cfg.nc, nm=cfg.rays, strides=cfg.strides, device=dev)
    crit._shapes = cfg.level_shapes
    packed, cap = crit.pack_targets(batch, cfg.batch, (320, 320))
    ref = po.pack_targets(batch, cfg.batch, (320, 320))
    assert packed.shape == ref.shape
    assert rel_err(packed.cpu(), ref, floor=1e-3) < 1e-6
    assert torch.equal(packed.cpu()[..., 0], ref[..., 0])
    m = po.in_box_mask((po.make_anchors(cfg.level_shapes, cfg.strides)[0] * po.make_anchors(cfg.level_shapes, cfg.strides)[1]),
                       ref[..., 1:5])
    assert int(m.sum()) <= cap                                       # host bound really is an upper bound


def test_capacity_overflow_is_loud():
    """A workspace sized for too few candidates must not produce a plausible-looking loss."""
    from ycr_b200.loss import v8SegmentationLoss
    from ycr_b200 import synth
    dev = _dev()
    cfg = synth.PathConfig("o", 2, 4, 160, nc=10)
    batch = synth.make_gts(cfg, 11)
    feats = [f.to(dev) for f in synth.make_feats(cfg, 11)]
    crit = v8SegmentationLoss(nc=cfg.nc, nm=cfg.rays, strides=cfg.strides, device=dev)
    crit._shapes = cfg.level_shapes
    packed, cap = crit.pack_targets(batch, cfg.batch, (160, 160))
    total, items = crit.call_packed(feats, packed, cap)
    assert torch.isfinite(total)
    total, items = crit.call_packed(feats, packed, 8)
    assert torch.isnan(total) and bool(torch.isnan(items).all())


def test_positive_targets_gather_equals_resweep():
    """The positives' ray targets come either from the rows K1 stored (default) or from a second sweep
    (when the store would not fit the budget): both paths must agree bit for bit."""
    import os
    from ycr_b200.loss import v8SegmentationLoss
    from ycr_b200.tal import TaskAlignedAssigner
    from ycr_b200 import synth
    dev = _dev()
    g = load_golden("train_s320_ragged")
    cfg, feats, batch = train_inputs(g)
    cpu, gpu, shapes = _assigner_inputs(cfg, feats, batch, dev)
    res = {}
    for mode, env in (("gather", None), ("resweep", "0")):
        if env is None:
            os.environ.pop("YCR_T_STORE_MAX_BYTES", None)
        else:
            os.environ["YCR_T_STORE_MAX_BYTES"] = env
        try:
            asg = TaskAlignedAssigner(topk=10, num_classes=cfg.nc, alpha=0.5, beta=4.0)
            out = asg(gpu["scores"], gpu["rays"], gpu["anc"], gpu["gl"], gpu["gb"], gpu["mask_gt"], gpu["gc"],
                      gpu["st"], gpu["ss"], 0, None, grid=(shapes, list(cfg.strides)))
            crit = v8SegmentationLoss(nc=cfg.nc, nm=cfg.rays, strides=cfg.strides, device=dev)
            fg = [f.to(dev).requires_grad_(True) for f in feats]
            total, items = crit((fg, 5, 2), batch)
            total.backward()
            res[mode] = (out[5].clone(), out[6].clone(), total.detach().clone(), [f.grad.clone() for f in fg])
        finally:
            os.environ.pop("YCR_T_STORE_MAX_BYTES", None)
    a, b = res["gather"], res["resweep"]
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2])
    for x, y in zip(a[3], b[3]):
        assert torch.equal(x, y)
    assert rel_err(a[0].cpu(), g["asg_gt_dist"]) < TOL


def test_clustered_contour_bin_count_wraps():
    """300 of the 360 contour points sit on a 3 px segment: from almost every in-box anchor they fall into
    one angular bin, whose byte-sized point count wraps.  The kernel must notice (the counts no longer add up
    to 360) and settle the sparse rays by the exact scan; results are checked against the oracle."""
    from ycr_b200.tal import TaskAlignedAssigner
    dev = _dev()
    shapes = [(40, 40), (20, 20), (10, 10)]
    strides = [8, 16, 32]
    anc, st = po.make_anchors(shapes, strides)
    A = anc.shape[0]
    gen = torch.Generator().manual_seed(77)
    # square 60..260 traversed by 60 points, plus a dense 3 px run on its right edge
    t = torch.linspace(0, 1, 61)[:-1]
    sq = torch.cat([torch.stack([60 + 200 * t[:15] / t[15], torch.full((15,), 60.)], 1),
                    torch.stack([torch.full((15,), 260.), 60 + 200 * t[:15] / t[15]], 1),
                    torch.stack([260 - 200 * t[:15] / t[15], torch.full((15,), 260.)], 1),
                    torch.stack([torch.full((15,), 60.), 260 - 200 * t[:15] / t[15]], 1)], 0)
    dense = torch.stack([torch.full((300,), 260.) + 0.013 * torch.arange(300) / 300,
                         150.0 + 3.0 * torch.arange(300) / 300 + 0.0007], 1)
    contour = torch.cat([sq[:23], dense, sq[23:]], 0)
    assert contour.shape == (360, 2)
    contour = contour + 0.01 * torch.rand(360, 2, generator=gen)
    gc = contour.reshape(1, 1, 720)
    gb = torch.tensor([[[60., 60., 260.2, 260.]]])
    scores = torch.rand(1, A, 2, generator=gen) * 0.8 + 0.1
    prays = torch.rand(1, A, 36, generator=gen) * 150 + 20
    asg = TaskAlignedAssigner(topk=10, num_classes=2, alpha=0.5, beta=4.0)
    asg.debug_metrics = True
    out = asg(scores.to(dev), prays.to(dev), (anc * st).to(dev), torch.zeros(1, 1, 1, device=dev), gb.to(dev),
              torch.ones(1, 1, 1, device=dev), gc.to(dev), st.to(dev), None, 0, None, grid=(shapes, strides))
    mp, gd = out[3].cpu(), out[5].cpu()
    pos = torch.nonzero(mp[0, 0]).flatten()
    assert pos.numel() == 10 and gd.shape == (10, 36)
    ref = po.polar_targets((anc * st)[pos], contour[None].expand(10, 360, 2), 36)
    ok = ~ref["ambiguous"]
    assert float(ok.float().mean()) > 0.9
    assert rel_err(gd[ok], ref["t"][ok]) < TOL
    # the dense overlaps of ALL candidates (every one swept the wrapped bin) against the oracle
    ov = asg.last_overlaps[0, 0].cpu()
    cand = torch.nonzero(po.in_box_mask(anc * st, gb)[0, 0]).flatten()
    sub = cand[torch.linspace(0, cand.numel() - 1, 60).long()]
    rt = po.polar_targets((anc * st)[sub], contour[None].expand(sub.numel(), 360, 2), 36)
    clean = ~rt["ambiguous"].any(1)
    ref_ov = po.polar_iou(prays[0, sub], rt["t"])
    assert int(clean.sum()) > 30
    assert rel_err(ov[sub][clean], ref_ov[clean]) < TOL


def test_assigner_grid_recovered_from_anchors():
    """Without `ss`/`imgsz` (and without the `grid` extension) the level shapes come from the anchors and the
    stride column themselves; same result as with the explicit grid."""
    from ycr_b200.tal import TaskAlignedAssigner
    dev = _dev()
    g = load_golden("train_s320_ragged")
    cfg, feats, batch = train_inputs(g)
    cpu, gpu, shapes = _assigner_inputs(cfg, feats, batch, dev)
    asg = TaskAlignedAssigner(topk=10, num_classes=cfg.nc, alpha=0.5, beta=4.0)
    a = asg(gpu["scores"], gpu["rays"], gpu["anc"], gpu["gl"], gpu["gb"], gpu["mask_gt"], gpu["gc"], gpu["st"],
            gpu["ss"], 0, None, grid=(shapes, list(cfg.strides)))
    b = asg(gpu["scores"], gpu["rays"], gpu["anc"], gpu["gl"], gpu["gb"], gpu["mask_gt"], gpu["gc"], gpu["st"],
            None, 0, None)
    for x, y in zip(a, b):
        assert torch.equal(x, y)
