"""GPU parity at BASELINE.json's own shapes (VERDICT r1 'next' item 1).

C2 (batch 64 @640, 20 GTs/img, 36 rays): the product runs the whole batch through the assigner and through the
fused loss; the oracle re-derives three sampled images completely (assignment is image-independent):
assigned GT, foreground mask, mask_pos bit-exact; polar targets, target scores, per-image loss numerators and
gradients at 1e-5.

C4 (batch 32 @1280, 200 GTs/img, 72 rays) is outside what the oracle can do per image in seconds (3e5 candidates
x 72 x 360 angle differences), so the check is split where the kernels split:
  * K1: dense Polar-IoU / align metric of sampled (image, GT) pairs against the oracle's polar targets;
  * K2/K3: per-GT top-k, multi-GT resolution and normalisation of the WHOLE batch re-derived with torch ops from
    the product's own dense metrics (utils/tal.py:1304-1338, :214-248, :1197-1202) - bit-exact;
  * K4: polar targets of the positives of sampled images against the oracle.
Also here: an anchor picked by several GTs whose overlaps are all zero (argmax -> GT 0, outside its box)."""
import pytest
import torch

from oracle import polar_oracle as po
from util import rel_err
from test_gpu_train import _assigner_inputs

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _run_assigner(cfg, gpu, shapes, debug=False):
    from ycr_b200.tal import TaskAlignedAssigner
    asg = TaskAlignedAssigner(topk=10, num_classes=cfg.nc, alpha=0.5, beta=4.0)
    asg.debug_metrics = debug
    out = asg(gpu["scores"], gpu["rays"], gpu["anc"], gpu["gl"], gpu["gb"], gpu["mask_gt"], gpu["gc"], gpu["st"],
              gpu["ss"], 0, None, grid=(shapes, list(cfg.strides)))
    return asg, out


def test_c2_full_batch_sampled_images_vs_oracle():
    from ycr_b200 import synth
    from ycr_b200.loss import v8SegmentationLoss
    from test_gpu_fullsize import _tss, _select_images
    dev = _dev()
    cfg = synth.CONFIGS["C2"]
    assert (cfg.batch, cfg.gts, cfg.imgsz, cfg.rays, cfg.nc) == (64, 20, 640, 36, 80)
    batch = synth.make_gts(cfg, 77)
    feats = synth.make_feats_near_gt(cfg, 77, batch)
    cpu, gpu, shapes = _assigner_inputs(cfg, feats, batch, dev)
    _, out = _run_assigner(cfg, gpu, shapes)
    tl, tb, ts, mp, tgi, gd, cen, fg = [t.cpu() for t in out]
    crit = v8SegmentationLoss(nc=cfg.nc, nm=cfg.rays, strides=cfg.strides, device=dev)
    fgpu = [f.to(dev).requires_grad_(True) for f in feats]
    total, items = crit((fgpu, 5, 2), batch)
    total.backward()
    tss_full = _tss(crit, [f.detach() for f in fgpu], batch)
    checked = 0
    for b in (3, 31, 63, 17, 48):
        sl = slice(b, b + 1)
        ref = po.assign(cpu["scores"][sl], cpu["rays"][sl], cpu["anc"], cpu["gl"][sl], cpu["gb"][sl],
                        cpu["mask_gt"][sl], cpu["gc"][sl])
        if not bool(ref["certain"][0]):
            continue
        assert torch.equal(tgi[sl], ref["target_gt_idx"])
        assert torch.equal(fg[sl], ref["fg_mask"])
        assert torch.equal(mp[sl], ref["mask_pos"])
        assert torch.equal(tl[sl], ref["target_labels"])
        nz = ref["target_scores"] != 0
        assert torch.equal(ts[sl] != 0, nz)
        assert rel_err(ts[sl][nz], ref["target_scores"][nz]) < 1e-5
        r0 = int(mp[:b].sum())
        rows = gd[r0:r0 + int(mp[sl].sum())]
        ok = ~ref["gt_dist_ambiguous"]
        assert rows.shape == ref["gt_dist"].shape
        assert rel_err(rows[ok], ref["gt_dist"][ok]) < TOL
        if bool(ref["gt_dist_ambiguous"].any()):
            continue   # the loss of this image is not pinned
        # fused loss: the oracle on the single image; numerators and gradients rescale with the normaliser
        one = po.seg_loss([f[sl] for f in feats], _select_images(batch, [b]), cfg.strides, cfg.nc, cfg.rays)
        scale = cfg.batch * one["target_scores_sum"] / tss_full
        for li in range(3):
            refg = one["grads"][li][0] * scale
            got = fgpu[li].grad[b].cpu()
            assert float((got - refg).abs().max()) <= TOL * float(refg.abs().max()), (b, li)
        checked += 1
        if checked == 3:
            break
    assert checked == 3, "fewer than three tie-free images among the sampled ones"


def _rederive(ov, al, in_box, valid, topk, eps):
    """select_topk_candidates + mask_pos + select_highest_overlaps + normaliser from dense (B,G,A) metrics, with torch
    ops on the device (utils/tal.py:1304-1338, :1216, :214-248, :1197-1202); ties lowest index (stable sort)."""
    B, G, A = ov.shape
    order = torch.sort(al, dim=2, descending=True, stable=True)[1][:, :, :topk]
    sel = torch.zeros_like(in_box)
    sel.scatter_(2, order, True)
    mask_pos = sel & in_box & valid[:, :, None]
    fgc = mask_pos.sum(1)
    multi = fgc > 1
    best = ov.argmax(1)
    one_hot = torch.zeros_like(mask_pos)
    one_hot.scatter_(1, best[:, None, :], True)
    mask_pos = torch.where(multi[:, None, :], one_hot, mask_pos)
    tgi = mask_pos.to(torch.uint8).argmax(1)
    fgm = mask_pos.any(1)
    alm = al * mask_pos
    pos_al = alm.amax(2, keepdim=True)
    pos_ov = (ov * mask_pos).amax(2, keepdim=True)
    norm = (alm * pos_ov / (pos_al + eps)).amax(1)
    return mask_pos, tgi, fgm, norm


def test_c4_full_batch_split_check():
    from ycr_b200 import synth
    dev = _dev()
    cfg = synth.CONFIGS["C4"]
    assert (cfg.batch, cfg.gts, cfg.imgsz, cfg.rays) == (32, 200, 1280, 72)
    batch = synth.make_gts(cfg, 88)
    feats = synth.make_feats(cfg, 88)
    cpu, gpu, shapes = _assigner_inputs(cfg, feats, batch, dev)
    asg, out = _run_assigner(cfg, gpu, shapes, debug=True)
    tl, tb, ts, mp, tgi, gd, cen, fg = out
    ov, al = asg.last_overlaps, asg.last_align_metric
    B, G, A = ov.shape
    # --- K1 on sampled (image, GT) pairs: every candidate's Polar-IoU inside the oracle's envelope ---
    anc = cpu["anc"]
    n_pairs = n_amb = n_rays = 0
    for b, g in [(0, 0), (5, 17), (13, 199), (21, 64), (31, 120), (9, 3)]:
        box = cpu["gb"][b, g]
        inb = po.in_box_mask(anc, box[None, None])[0, 0]
        ai = torch.nonzero(inb).flatten()
        if ai.numel() == 0:
            continue
        contour = cpu["gc"][b, g].view(1, 360, 2).expand(ai.numel(), -1, -1)
        pt = po.polar_targets(anc[ai], contour, 72)
        pr = cpu["rays"][b, ai]
        lo = torch.minimum(pr, pt["t_lo"]).clamp(min=po.FLOOR).sum(-1) / torch.maximum(pr, pt["t_hi"]).sum(-1)
        hi = torch.minimum(pr, pt["t_hi"]).clamp(min=po.FLOOR).sum(-1) / torch.maximum(pr, pt["t_lo"]).sum(-1)
        got = ov[b, g, ai.to(dev)].cpu()
        assert bool(((got >= lo * (1 - TOL)) & (got <= hi * (1 + TOL))).all()), (b, g)
        sure = lo == hi
        assert rel_err(got[sure], po.polar_iou(pt["t"], pr)[sure]) < TOL
        off = torch.ones(A, dtype=torch.bool)
        off[ai] = False
        assert float(ov[b, g, off.to(dev)].abs().max()) == 0.0          # exactly the in-box anchors are candidates
        n_pairs += 1
        n_amb += int(pt["ambiguous"].sum())
        n_rays += pt["ambiguous"].numel()
    assert n_pairs >= 4
    assert n_amb / n_rays < 0.01
    # --- K2/K3 on the whole batch, re-derived from the product's own dense metrics ---
    in_box = ov != 0                                   # candidates have a positive Polar-IoU (sum of floors > 0)
    valid = gpu["mask_gt"][:, :, 0] > 0
    mask_pos_r, tgi_r, fg_r, norm_r = _rederive(ov, al, in_box, valid, 10, 1e-9)
    # a zero-metric in-box candidate as top-k filler would make the comparison depend on tie order: none here
    assert bool((al[in_box] > 0).all())
    assert torch.equal(mp, mask_pos_r)
    assert torch.equal(fg, fg_r)
    assert torch.equal(tgi, tgi_r)
    tsum = ts.sum(-1)
    assert torch.equal(tsum != 0, fg_r & (norm_r != 0))
    nz = tsum != 0
    assert rel_err(tsum[nz].cpu(), norm_r[nz].cpu()) < TOL
    lab = gpu["gl"][:, :, 0].long().gather(1, tgi_r)
    assert torch.equal(tl, lab.clamp(min=0))
    # --- K4 on sampled images: polar targets of the positives, (b,g,a) order ---
    mpc = mp.cpu()
    rows_before = torch.cumsum(mpc.view(B, -1).sum(1), 0)
    gdc = gd.cpu()
    for b in (2, 30):
        pg, pa = torch.nonzero(mpc[b], as_tuple=True)
        contour = cpu["gc"][b].view(G, 360, 2)[pg]
        pt = po.polar_targets(anc[pa], contour, 72)
        r0 = int(rows_before[b - 1]) if b else 0
        rows = gdc[r0:r0 + pg.numel()]
        ok = ~pt["ambiguous"]
        assert rel_err(rows[ok], pt["t"][ok]) < TOL
        assert float(ok.float().mean()) > 0.99


def test_multi_picked_anchor_with_all_overlaps_zero_goes_to_gt0_with_real_geometry():
    """Two small GTs over the image corner pick anchor 0 as a zero-metric top-k filler (its predictions are +inf, so its
    Polar-IoU with both is 0); argmax over the all-zero overlap column is GT 0 (utils/tal.py:231), whose box does not
    hold the anchor; the reference still computes the polygon->polar targets of that (GT 0, anchor 0) pair
    (utils/tal.py:1172-1193)."""
    from ycr_b200 import synth
    dev = _dev()
    cfg = synth.PathConfig("z", 1, 3, 64, nc=3)
    t = torch.linspace(0, 2 * torch.pi, 361)[:-1]

    def poly(cx, cy, r):
        p = torch.stack([cx + r * torch.cos(t), cy + r * torch.sin(t)], 1).numpy() / 64.0
        return torch.from_numpy(synth.resample_closed(p))
    segs = torch.stack([poly(44.3, 43.1, 12.2), poly(6.1, 5.9, 5.3), poly(5.2, 6.3, 4.6)])
    lo, hi = segs.min(1)[0], segs.max(1)[0]
    batch = {"batch_idx": torch.zeros(3), "cls": torch.tensor([[0.], [1.], [2.]]),
             "bboxes": torch.cat(((lo + hi) / 2, hi - lo), 1), "segments": [segs]}
    feats = synth.make_feats(cfg, 5)
    feats[0][0, :cfg.rays, 0, 0] = float("inf")        # anchor 0 = cell (0,0) of the stride-8 level, centre (4,4)
    cpu, gpu, shapes = _assigner_inputs(cfg, feats, batch, dev)
    _, out = _run_assigner(cfg, gpu, shapes)
    tl, tb, ts, mp, tgi, gd, cen, fg = [x.cpu() for x in out]
    ref = po.assign(cpu["scores"], cpu["rays"], cpu["anc"], cpu["gl"], cpu["gb"], cpu["mask_gt"], cpu["gc"])
    assert bool(ref["mask_pos"][0, 0, 0]) and not bool(po.in_box_mask(cpu["anc"][:1], cpu["gb"][:, :1])[0, 0, 0]), \
        "the constructed case must put anchor 0 on GT 0 from outside its box"
    assert torch.equal(mp, ref["mask_pos"])
    assert torch.equal(tgi, ref["target_gt_idx"])
    assert torch.equal(fg, ref["fg_mask"])
    ok = ~ref["gt_dist_ambiguous"]
    assert gd.shape == ref["gt_dist"].shape
    assert rel_err(gd[ok], ref["gt_dist"][ok]) < TOL
    assert rel_err(gd[0][ok[0]], ref["gt_dist"][0][ok[0]]) < TOL       # row 0 is the (GT 0, anchor 0) pair


def test_more_than_255_gts_per_image():
    """round 1 rejected G > 255.  The limit is now what the per-image resolution kernel can hold in shared
    memory (815 GTs at 640 px); 300 small instances in one image against the oracle."""
    from ycr_b200 import synth
    dev = _dev()
    cfg = synth.PathConfig("g300", 1, 300, 320, nc=5)
    batch = synth.make_gts(cfg, 61)
    batch["segments"] = [0.5 + (s - s.mean(1, keepdim=True)) * 0.35 + (s.mean(1, keepdim=True) - 0.5) for s in batch["segments"]]
    lo, hi = batch["segments"][0].min(1)[0], batch["segments"][0].max(1)[0]
    batch["bboxes"] = torch.cat(((lo + hi) / 2, hi - lo), 1)
    feats = synth.make_feats_near_gt(cfg, 61, batch)
    cpu, gpu, shapes = _assigner_inputs(cfg, feats, batch, dev)
    assert cpu["gb"].shape[1] == 300
    _, out = _run_assigner(cfg, gpu, shapes)
    tl, tb, ts, mp, tgi, gd, cen, fg = [t.cpu() for t in out]
    ref = po.assign(cpu["scores"], cpu["rays"], cpu["anc"], cpu["gl"], cpu["gb"], cpu["mask_gt"], cpu["gc"])
    if bool(ref["certain"].all()):
        assert torch.equal(mp, ref["mask_pos"]) and torch.equal(tgi, ref["target_gt_idx"]) and torch.equal(fg, ref["fg_mask"])
    else:   # a near-tie somewhere among 300 crowded GTs: the undisputed part must still agree
        assert float((mp != ref["mask_pos"]).float().mean()) < 1e-4
    ok = ~ref["gt_dist_ambiguous"]
    if gd.shape == ref["gt_dist"].shape:
        assert rel_err(gd[ok], ref["gt_dist"][ok]) < TOL
    # and beyond the limit the error names it
    from ycr_b200.tal import TaskAlignedAssigner
    asg = TaskAlignedAssigner(topk=10, num_classes=cfg.nc, alpha=0.5, beta=4.0)
    G_big = 1200
    with pytest.raises(ValueError, match="shared memory"):
        asg(gpu["scores"], gpu["rays"], gpu["anc"], torch.zeros(1, G_big, 1, device=dev), torch.zeros(1, G_big, 4, device=dev),
            torch.zeros(1, G_big, 1, device=dev), torch.zeros(1, G_big, 720, device=dev), gpu["st"], gpu["ss"], 0, None,
            grid=(shapes, list(cfg.strides)))


def test_assigner_grid_cache_distinguishes_transposed_images():
    """a 320x256 and a 256x320 image have the same anchor count per level; the level shapes must follow the
    image size, not a cache entry of the first call."""
    from ycr_b200 import synth
    from ycr_b200.tal import TaskAlignedAssigner
    dev = _dev()
    asg = TaskAlignedAssigner(topk=10, num_classes=4, alpha=0.5, beta=4.0)
    g = torch.Generator().manual_seed(3)
    for (H, W) in ((320, 256), (256, 320)):
        shapes = [(H // s, W // s) for s in (8, 16, 32)]
        anc, st = po.make_anchors(shapes, (8, 16, 32))
        A = anc.shape[0]
        t = torch.linspace(0, 2 * torch.pi, 361)[:-1]
        cx, cy, r = 0.55 * W, 0.45 * H, 0.2 * min(H, W)
        poly = torch.stack([cx + r * (1 + 0.2 * torch.sin(3 * t)) * torch.cos(t), cy + r * torch.sin(t)], 1)
        seg = torch.from_numpy(synth.resample_closed((poly / torch.tensor([W, H])).numpy())) * torch.tensor([W, H])
        box = torch.cat((seg.min(0)[0], seg.max(0)[0]))[None, None]
        scores = torch.rand(1, A, 4, generator=g) * 0.5 + 0.1
        rays = (torch.rand(1, A, 36, generator=g) * 3 + 1) * st
        gl = torch.tensor([[[2.0]]])
        mask = torch.ones(1, 1, 1)
        gc = seg.reshape(1, 1, 720)
        ss = [torch.full((h * w, 1), float(s)) for (h, w), s in zip(shapes, (8, 16, 32))]
        out = asg(scores.to(dev), rays.to(dev), (anc * st).to(dev), gl.to(dev), box.to(dev), mask.to(dev), gc.to(dev),
                  st.to(dev), [s.to(dev) for s in ss], 0, torch.tensor([float(H), float(W)], device=dev))
        ref = po.assign(scores, rays, anc * st, gl, box, mask, gc)
        assert torch.equal(out[3].cpu(), ref["mask_pos"]), (H, W)
        assert torch.equal(out[7].cpu(), ref["fg_mask"]), (H, W)
