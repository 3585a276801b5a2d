"""GPU, BASELINE.json's full sizes (C2 training batch 64 @640 / 20 GTs, C3 inference batch 256 @640):
the oracle cannot run these in seconds, so parity is checked through size-independent properties of the
domain — image independence (permuting / splitting the batch), additivity of the loss numerators,
structural invariants of the assignment, and the defining properties of decode and greedy NMS."""
import pytest
import torch

from oracle import polar_oracle as po

pytestmark = pytest.mark.gpu


def _dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _c2(dev, seed=301):
    from ycr_b200 import synth
    cfg = synth.CONFIGS["C2"]
    batch = synth.make_gts(cfg, seed)
    small = synth.PathConfig("g", 16, cfg.gts, cfg.imgsz, rays=cfg.rays, nc=cfg.nc)
    sub = {"batch_idx": batch["batch_idx"][batch["batch_idx"] < 16], "cls": batch["cls"][batch["batch_idx"] < 16],
           "bboxes": batch["bboxes"][batch["batch_idx"] < 16], "segments": batch["segments"][:16]}
    f16 = synth.make_feats_near_gt(small, seed, sub)
    feats = [torch.cat([f.roll(k, 0) for k in range(4)], 0).contiguous().to(dev) for f in f16]
    return cfg, batch, feats


def _select_images(batch, idx):
    """Sub-batch made of the images `idx` (in that order), image index re-based."""
    bi = batch["batch_idx"].long()
    rows, new_bi = [], []
    for new, old in enumerate(idx):
        r = torch.nonzero(bi == old).flatten()
        rows.append(r)
        new_bi.append(torch.full((r.numel(),), float(new)))
    rows = torch.cat(rows)
    return {"batch_idx": torch.cat(new_bi), "cls": batch["cls"][rows], "bboxes": batch["bboxes"][rows],
            "segments": [batch["segments"][o] for o in idx]}


def test_c2_loss_is_image_independent_and_additive():
    from ycr_b200.loss import v8SegmentationLoss
    dev = _dev()
    cfg, batch, feats = _c2(dev)
    crit = v8SegmentationLoss(nc=cfg.nc, nm=cfg.rays, strides=cfg.strides, device=dev)

    def run(fs, b):
        fl = [f.clone().requires_grad_(True) for f in fs]
        total, items = crit((fl, 5, 2), b)
        total.backward()
        return total.detach(), items, [f.grad for f in fl]

    B = cfg.batch
    total, items, grads = run(feats, batch)
    assert torch.isfinite(total) and bool(torch.isfinite(items).all())
    # (1) permutation of the images permutes the gradients and leaves the loss (almost) unchanged
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(5)).tolist()
    total_p, items_p, grads_p = run([f[perm] for f in feats], _select_images(batch, perm))
    assert float((items_p - items).abs().max()) <= 2e-6 * float(items.abs().max())
    for g, gp in zip(grads, grads_p):                         # (the normaliser is summed in another order: 1 ulp)
        assert float((g[perm] - gp).abs().max()) <= 2e-6 * float(g.abs().max())
    # (2) splitting the batch: loss numerators add up.  items = gain * sum / tss  =>  sum = items * tss / gain
    halves = []
    for lo in (0, B // 2):
        idx = list(range(lo, lo + B // 2))
        t, it, gr = run([f[idx] for f in feats], _select_images(batch, idx))
        halves.append((it, gr, idx))
    tss_full = _tss(crit, feats, batch)
    tss_h = [_tss(crit, [f[h[2]] for f in feats], _select_images(batch, h[2])) for h in halves]
    assert abs(sum(tss_h) - tss_full) <= 1e-5 * tss_full
    for k in range(2):
        num_full = float(items[k]) * tss_full
        num_half = sum(float(h[0][k]) * t for h, t in zip(halves, tss_h))
        assert abs(num_full - num_half) <= 2e-5 * abs(num_full)
    # gradients: full-batch grad of image b = half-batch grad * (tss_half / tss_full) * (B / (B/2))
    for li in range(3):
        for h, t in zip(halves, tss_h):
            scale = (t / tss_full) * 2.0
            ref = h[1][li] * scale
            got = grads[li][h[2]]
            assert float((got - ref).abs().max()) <= 2e-5 * float(ref.abs().max())


def _tss(crit, feats, batch):
    """target_scores_sum of a batch, read from the library's loss_out[3] through the autograd function."""
    from ycr_b200.loss import _SegLossFn
    from ycr_b200.tal import gt_struct
    B = feats[0].shape[0]
    crit._shapes = [tuple(f.shape[2:]) for f in feats]
    packed, cap = crit.pack_targets(batch, B, (feats[0].shape[2] * 8, feats[0].shape[3] * 8))
    gl, gb, gc = packed.split((1, 4, 720), 2)
    gt, keep = gt_struct(gl, gb, gc, None)
    total, out = _SegLossFn.apply(crit, gt, cap, *feats)
    return float(out[3])


def test_c2_assignment_invariants():
    from ycr_b200.tal import TaskAlignedAssigner
    from ycr_b200 import synth
    dev = _dev()
    cfg, batch, feats = _c2(dev, seed=302)
    B, no = cfg.batch, cfg.rays + cfg.nc
    cat = torch.cat([f.view(B, no, -1) for f in feats], 2)
    rays, logits = cat.split((cfg.rays, cfg.nc), 1)
    anc, st = po.make_anchors(cfg.level_shapes, cfg.strides)
    anc, st = anc.to(dev), st.to(dev)
    t = po.pack_targets(batch, B, (cfg.imgsz, cfg.imgsz)).to(dev)
    gl, gb, gc = t.split((1, 4, 720), 2)
    mg = (gb.sum(2, keepdim=True) > 0).float()
    scores = logits.permute(0, 2, 1).contiguous().sigmoid()
    prays = rays.permute(0, 2, 1).contiguous() * st
    asg = TaskAlignedAssigner(topk=10, num_classes=cfg.nc, alpha=0.5, beta=4.0)
    tl, tb, ts, mp, tgi, gd, cen, fg = asg(scores, prays, anc * st, gl, gb, mg, gc, st, None, 0, None,
                                             grid=(cfg.level_shapes, list(cfg.strides)))
    A = anc.shape[0]
    assert mp.shape == (B, cfg.gts, A) and gd.shape == (int(mp.sum()), cfg.rays)
    assert torch.equal(mp.sum(1) > 0, fg) and int(mp.sum(1).max()) <= 1           # one GT per anchor at most
    b_i, g_i, a_i = torch.nonzero(mp, as_tuple=True)
    assert torch.equal(tgi[b_i, a_i], g_i)
    assert bool((tgi[~fg] == 0).all())
    # every positive anchor lies strictly inside the box of its GT (candidates are in-box anchors)
    ap = (anc * st)[a_i]
    bx = gb[b_i, g_i]
    assert bool(((ap[:, 0] > bx[:, 0]) & (ap[:, 0] < bx[:, 2]) & (ap[:, 1] > bx[:, 1]) & (ap[:, 1] < bx[:, 3])).all())
    # at least one and at most a few times topk positives per valid GT; labels / boxes gathered from that GT
    per_gt = mp.sum(2)
    assert int(per_gt.max()) <= 4 * 10 and float((per_gt > 0).float().mean()) > 0.9
    assert torch.equal(tl[b_i, a_i], gl[b_i, g_i, 0].long()) and torch.equal(tb[b_i, a_i], bx)
    # target scores: one non-zero per positive, at the GT's class, in (0, 1]; zero rows elsewhere
    nz = ts != 0
    assert int(nz.sum()) == int(fg.sum()) and bool((ts[b_i, a_i, tl[b_i, a_i]] > 0).all())
    assert float(ts.max()) <= 1.0 + 1e-6 and not bool(nz[~fg].any())
    # per GT the best positive's normalised score equals that GT's best overlap (<= 1)
    assert bool((gd >= 1e-6).all()) and bool((cen > 0).all()) and bool((cen <= 1 + 1e-6).all())
    # (b,g,a)-lexicographic row order of gt_dist: row r belongs to the r-th nonzero of mask_pos -> spot check
    # three rows against the oracle's polar targets
    contour = gc.view(B, cfg.gts, 360, 2)
    pick = torch.linspace(0, gd.shape[0] - 1, 3).long()
    ref = po.polar_targets(ap[pick].cpu(), contour[b_i[pick], g_i[pick]].cpu(), cfg.rays)
    ok = ~ref["ambiguous"]
    assert float(((gd[pick].cpu() - ref["t"]).abs() / ref["t"])[ok].max()) < 1e-5


def test_c3_decode_and_nms_properties():
    from ycr_b200.head import decode
    from ycr_b200.ops import non_max_suppression
    from ycr_b200 import synth
    dev = _dev()
    cfg = synth.CONFIGS["C3"]
    small = synth.make_feats(synth.PathConfig("gi", 16, 0, cfg.imgsz, rays=cfg.rays, nc=cfg.nc), 303)
    feats = [torch.cat([f.roll(k, 0) for k in range(cfg.batch // 16)], 0).contiguous().to(dev) for f in small]
    R, nc = cfg.rays, cfg.nc
    out = decode(feats, cfg.strides, nc, R)
    B, CH, A = out.shape
    assert (B, CH, A) == (256, 4 + nc + 3 * R, 8400)
    x, y, v = out[:, 4 + nc:4 + nc + R], out[:, 4 + nc + R:4 + nc + 2 * R], out[:, 4 + nc + 2 * R:]
    assert torch.equal(out[:, 0], x.min(1)[0]) and torch.equal(out[:, 2], x.max(1)[0])     # box = extents of the contour
    assert torch.equal(out[:, 1], y.min(1)[0]) and torch.equal(out[:, 3], y.max(1)[0])
    assert bool(((v == 0) | (v == 1)).all())
    cls = out[:, 4:4 + nc]
    assert bool((cls > 0).all()) and bool((cls < 1).all())
    dets = non_max_suppression(out, 0.25, 0.7, nc=nc, max_det=300)
    assert len(dets) == B
    thr = 0.7
    for b in (0, 37, 255):
        d = dets[b]
        assert d.shape[1] == 6 + 3 * R and d.shape[0] <= 300
        assert bool((d[1:, 4] <= d[:-1, 4]).all())                       # descending score
        assert bool((d[:, 4] > 0.25).all())
        # kept boxes of one class never overlap by more than the threshold (on the class-offset boxes)
        boxes = d[:, :4] + d[:, 5:6] * 7680
        area = (boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1])
        lt = torch.max(boxes[:, None, :2], boxes[None, :, :2])
        rb = torch.min(boxes[:, None, 2:], boxes[None, :, 2:])
        inter = (rb - lt).clamp(min=0).prod(2)
        iou = inter / (area[:, None] + area[None] - inter)
        iou.fill_diagonal_(0)
        assert float(iou.max()) <= thr
        # rows are gathered verbatim from the prediction tensor
        conf, j = cls[b].max(0)
        a_idx = torch.nonzero((out[b, 0][None] == d[:, 0:1]) & (out[b, 1][None] == d[:, 1:2]) &
                              (conf[None] == d[:, 4:5]))
        assert a_idx.shape[0] >= d.shape[0]
    # idempotence: suppressing the kept set again keeps everything (greedy NMS fixed point)
    d = dets[0]
    pred2 = torch.zeros(1, CH, max(d.shape[0], 1), device=dev)
    pred2[0, :4] = d[:, :4].T
    pred2[0, 4 + d[:, 5].long(), torch.arange(d.shape[0], device=dev)] = d[:, 4]
    pred2[0, 4 + nc:] = d[:, 6:].T
    again = non_max_suppression(pred2, 0.25, 0.7, nc=nc, max_det=300)[0]
    assert again.shape[0] == d.shape[0] and torch.equal(again[:, :6], d[:, :6])


def test_c2_loss_bit_identical_run_to_run():
    """Full-size batch: chunks are handed out dynamically and the two warps of a block share their queues, yet
    every value must come out bit-identical on every run (no order-dependent reduction anywhere)."""
    from ycr_b200.loss import v8SegmentationLoss
    dev = _dev()
    cfg, batch, feats = _c2(dev, seed=305)
    crit = v8SegmentationLoss(nc=cfg.nc, nm=cfg.rays, strides=cfg.strides, device=dev)
    ref = None
    for _ in range(4):
        fl = [f.clone().requires_grad_(True) for f in feats]
        total, items = crit((fl, 5, 2), batch)
        total.backward()
        cur = (total.detach().clone(), items.clone(), [f.grad.clone() for f in fl])
        if ref is None:
            ref = cur
            continue
        assert torch.equal(ref[0], cur[0]) and torch.equal(ref[1], cur[1])
        for a, b in zip(ref[2], cur[2]):
            assert torch.equal(a, b)


def test_fresh_criteria_back_to_back_calls_do_not_disturb_each_other():
    """The GT rows travel on a side stream into buffers a new criterion allocates on its first two calls; the caching
    allocator may hand it a block that kernels still queued on the compute stream are using (the packed GT tensor of the
    call before has the same size).  Calls issued back to back, without a synchronisation in between, from criteria
    created over and over must all return the same bits."""
    from ycr_b200.loss import v8SegmentationLoss
    dev = _dev()
    cfg, batch, feats = _c2(dev, seed=306)
    want = None
    for _ in range(6):
        crit = v8SegmentationLoss(nc=cfg.nc, nm=cfg.rays, strides=cfg.strides, device=dev)
        outs = []
        for _ in range(3):                      # the host runs ahead of the device here
            fl = [f.clone().requires_grad_(True) for f in feats]
            total, items = crit((fl, 5, 2), batch)
            total.backward()
            outs.append((total.detach(), items, fl[0].grad))
            del fl
        got = [(float(t), i.tolist(), float(g.double().sum())) for t, i, g in outs]
        if want is None:
            want = got[0]
        for g in got:
            assert g == want
        del crit
