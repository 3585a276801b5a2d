"""GPU parity tests of the inference path: Segment decode and non_max_suppression, through the
C-ABI library via the host mirror, against reference-generated golden vectors and the oracle."""
import numpy as np
import pytest
import torch

from util import load_golden, infer_inputs, split_rows
from oracle import polar_oracle as po

pytestmark = pytest.mark.gpu

NMS_CASES = (("best", dict(conf_thres=0.25, iou_thres=0.7, multi_label=False)),
             ("multi", dict(conf_thres=0.25, iou_thres=0.7, multi_label=True)),
             ("agn", dict(conf_thres=0.4, iou_thres=0.5, agnostic=True, max_det=50)))


def _dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


@pytest.mark.parametrize("name", ["infer_s160", "infer_s320", "infer_c3small"])
def test_decode_matches_reference(name):
    from ycr_b200.head import decode
    dev = _dev()
    g = load_golden(name)
    cfg, feats = infer_inputs(g)
    out = decode([f.to(dev) for f in feats], cfg.strides, cfg.nc, cfg.rays).cpu()
    ref = po.decode(feats, cfg.strides, cfg.nc, cfg.rays)
    if "allpred" in g:
        assert np.array_equal(ref.numpy(), g["allpred"])
    assert out.shape == ref.shape
    # fp32 tolerance 1e-5 relative (+1e-5 px absolute for coordinates that cancel to ~0)
    assert bool(((out - ref).abs() <= 1e-5 * ref.abs() + 1e-5).all())
    R, nc = cfg.rays, cfg.nc
    assert torch.equal(out[:, 4 + nc + 2 * R:], ref[:, 4 + nc + 2 * R:])      # validity flags: exact
    frac_exact = float((out == ref).float().mean())
    assert frac_exact > 0.5, frac_exact        # sigmoid/sincos differ from torch-CPU by <= 1 ulp


@pytest.mark.parametrize("name", ["infer_s160", "infer_s320", "infer_c3small"])
def test_nms_matches_reference_keep_lists(name):
    """NMS in isolation: feed the reference's own decoded tensor, expect identical rows."""
    from ycr_b200.ops import non_max_suppression
    dev = _dev()
    g = load_golden(name)
    cfg, feats = infer_inputs(g)
    allpred = po.decode(feats, cfg.strides, cfg.nc, cfg.rays)
    for tag, kw in NMS_CASES:
        dets = non_max_suppression(allpred.to(dev), nc=cfg.nc, **kw)
        assert [d.shape[0] for d in dets] == g[f"nms_{tag}_counts"].tolist(), tag
        ref = split_rows(g[f"nms_{tag}_rows"], g[f"nms_{tag}_counts"])
        for d, r in zip(dets, ref):
            assert np.array_equal(d.cpu().numpy(), r), tag


def test_decode_then_nms_pipeline():
    """End to end on the product's own decode output: same keep-lists as the oracle pipeline."""
    from ycr_b200.head import decode
    from ycr_b200.ops import non_max_suppression
    dev = _dev()
    g = load_golden("infer_s320")
    cfg, feats = infer_inputs(g)
    out = decode([f.to(dev) for f in feats], cfg.strides, cfg.nc, cfg.rays)
    dets = non_max_suppression(out, 0.25, 0.7, nc=cfg.nc)
    ref, margin = po.nms(po.decode(feats, cfg.strides, cfg.nc, cfg.rays), 0.25, 0.7, nc=cfg.nc)
    assert margin > 1e-5
    assert [d.shape[0] for d in dets] == [r.shape[0] for r in ref]
    for d, r in zip(dets, ref):
        assert torch.equal(d[:, 5].cpu(), r[:, 5])
        assert bool(((d.cpu() - r).abs() <= 1e-5 * r.abs() + 1e-5).all())


def _random_pred(rng, B, A, nc, nm, n_hot, dup=False):
    pred = np.zeros((B, 4 + nc + nm, A), np.float32)
    xy = rng.uniform(0, 600, size=(B, 2, A)).astype(np.float32)
    wh = rng.uniform(8, 160, size=(B, 2, A)).astype(np.float32)
    pred[:, 0:2] = xy
    pred[:, 2:4] = xy + wh
    pred[:, 4:4 + nc] = rng.uniform(0, 0.2, size=(B, nc, A)).astype(np.float32)
    for b in range(B):
        hot = rng.choice(A, size=min(A, n_hot), replace=False)
        cls = rng.integers(0, nc, size=hot.size)
        pred[b, 4 + cls, hot] = rng.uniform(0.3, 1.0, size=hot.size).astype(np.float32)
        # clusters of near-duplicates so suppression actually happens
        for k in range(0, hot.size - 1, 2):
            pred[b, 0:4, hot[k + 1]] = pred[b, 0:4, hot[k]] + rng.uniform(-6, 6, size=4).astype(np.float32)
            pred[b, 4:4 + nc, hot[k + 1]] = 0.05
            pred[b, 4 + cls[k], hot[k + 1]] = rng.uniform(0.3, 1.0)
        if dup and hot.size > 4:
            pred[b, :, hot[3]] = pred[b, :, hot[2]]          # identical box and equal score: input order wins
    pred[:, 4 + nc:] = rng.standard_normal((B, nm, A)).astype(np.float32)
    return torch.from_numpy(pred)


@pytest.mark.parametrize("A,n_hot,kw", [
    (2100, 300, dict(conf_thres=0.25, iou_thres=0.7)),
    (2100, 0, dict(conf_thres=0.25, iou_thres=0.7)),                       # nothing passes: empty outputs
    (8400, 2000, dict(conf_thres=0.25, iou_thres=0.45, max_det=300)),
    (8400, 6000, dict(conf_thres=0.25, iou_thres=0.6, max_det=1000)),       # > 4096 candidates: global sort
    (8400, 1500, dict(conf_thres=0.25, iou_thres=0.5, multi_label=True)),
    (8400, 1500, dict(conf_thres=0.25, iou_thres=0.5, classes=[1, 3, 5])),
    (8400, 1500, dict(conf_thres=0.25, iou_thres=0.5, classes=[0, 2], multi_label=True)),
    (8400, 3000, dict(conf_thres=0.25, iou_thres=0.5, max_nms=1000)),
    (8400, 1500, dict(conf_thres=0.25, iou_thres=0.5, agnostic=True)),
    # the validator's regime: conf 0.001 with multi-label lets ~67 k (anchor, class) pairs per image through,
    # more than twice max_nms - the kernel selects the best max_nms before sorting
    (8400, 1501, dict(conf_thres=0.001, iou_thres=0.6, multi_label=True, max_det=300, _batch=1)),
    (8400, 1502, dict(conf_thres=0.001, iou_thres=0.6, multi_label=True, max_nms=5000, max_det=100, _batch=1)),
])
def test_nms_random_vs_oracle(A, n_hot, kw):
    from ycr_b200.ops import non_max_suppression
    dev = _dev()
    rng = np.random.default_rng(A + n_hot)
    nc, nm = 8, 12
    kw = dict(kw)
    pred = _random_pred(rng, kw.pop("_batch", 3), A, nc, nm, n_hot, dup=True)
    ref, margin = po.nms(pred, nc=nc, **kw)
    assert margin > 1e-6
    dets = non_max_suppression(pred.to(dev), nc=nc, **kw)
    assert [d.shape[0] for d in dets] == [r.shape[0] for r in ref]
    for d, r in zip(dets, ref):
        assert d.shape[1] == 6 + nm
        assert torch.equal(d.cpu(), r)


@pytest.mark.parametrize("multi", [False, True], ids=["single_label_tranche", "multi_label_score_histogram"])
def test_nms_leading_entries_run_out_and_the_image_is_redone(multi):
    """More than 1024 candidates whose best-scored ones are near-duplicates of a few boxes: the entries the kernels look
    at first (the sorted leading tranche; with multi-label over a large class space the entries above a per-image
    score-bin limit) yield far fewer than max_det survivors, so image 0 has to be redone over everything; image 1
    reaches max_det at once; image 2 has few candidates.  Same rows as the oracle."""
    from ycr_b200.ops import non_max_suppression
    dev = _dev()
    rng = np.random.default_rng(4242)
    nc, nm, A, B = (8, 6, 8400, 3) if multi else (3, 6, 4200, 3)
    pred = np.zeros((B, 4 + nc + nm, A), dtype=np.float32)
    if multi:   # every (anchor, class) pair passes conf 0.001; the background anchors are random boxes
        xy = rng.uniform(0, 600, size=(B, 2, A)).astype(np.float32)
        pred[:, 0:2] = xy
        pred[:, 2:4] = xy + rng.uniform(8, 60, size=(B, 2, A)).astype(np.float32)
        pred[:, 4:4 + nc] = rng.uniform(0.002, 0.2, size=(B, nc, A)).astype(np.float32)
    else:
        pred[:, 4:4 + nc] = 0.01

    def put(b, idx, xyxy, cls, score):
        pred[b, 0:4, idx] = xyxy
        pred[b, 4 + cls, idx] = score

    scores = rng.permutation(np.linspace(0.3, 0.999, 3 * 4200).astype(np.float32))   # all distinct
    si = iter(scores)
    # image 0: 1500 high-scored boxes in 6 tight clusters, then 900 well separated lower-scored boxes
    centres = rng.uniform(100, 540, size=(6, 2))
    hi = sorted([float(next(si)) for _ in range(2400)], reverse=True)
    for k in range(1500):
        c = centres[k % 6] + rng.uniform(-1.0, 1.0, size=2)
        put(0, k, [c[0] - 40, c[1] - 40, c[0] + 40, c[1] + 40], 0, hi[k])
    for k in range(900):
        gx, gy = (k % 30) * 21 + 5, (k // 30) * 21 + 5
        put(0, 1500 + k, [gx, gy, gx + 9, gy + 9], 1, hi[1500 + k])
    # image 1: 2000 separated boxes (max_det reached at once)
    for k in range(2000):
        gx, gy = (k % 50) * 12 + 3, (k // 50) * 12 + 3
        put(1, k, [gx, gy, gx + 8, gy + 8], 2, float(next(si)))
    # image 2: 700 strong candidates only
    for k in range(700):
        gx, gy = (k % 30) * 20 + 3, (k // 30) * 20 + 3
        put(2, k, [gx, gy, gx + 15, gy + 15], k % nc, float(next(si)))
    pred[:, 4 + nc:] = rng.standard_normal((B, nm, A)).astype(np.float32)
    pred = torch.from_numpy(pred)
    kw = dict(conf_thres=0.001, iou_thres=0.6, max_det=300, multi_label=True) if multi else \
        dict(conf_thres=0.25, iou_thres=0.6, max_det=300)
    ref, margin = po.nms(pred, nc=nc, **kw)
    assert margin > 1e-6
    assert ref[0].shape[0] == 300 and int((ref[0][:, 5] == 0).sum()) <= 12     # a handful of cluster survivors, then class 1
    dets = non_max_suppression(pred.to(dev), nc=nc, **kw)
    assert [d.shape[0] for d in dets] == [r.shape[0] for r in ref]
    for d, r in zip(dets, ref):
        assert torch.equal(d.cpu(), r)


def test_detect_packed_output_is_the_concatenated_list():
    from ycr_b200.ops import detect
    from ycr_b200 import synth
    dev = _dev()
    cfg = synth.PathConfig("dp", 3, 0, 320, nc=12)
    feats = [f.to(dev) for f in synth.make_feats(cfg, 19)]
    lst = detect(feats, cfg.strides, cfg.nc, cfg.rays, 0.25, 0.7)
    rows, counts = detect(feats, cfg.strides, cfg.nc, cfg.rays, 0.25, 0.7, packed=True)
    n = counts.tolist()
    assert n == [d.shape[0] for d in lst] and sum(n) > 0
    assert torch.equal(rows[:sum(n)], torch.cat(lst))


def test_nms_argument_errors():
    from ycr_b200.ops import non_max_suppression
    dev = _dev()
    p = torch.zeros(1, 4 + 3 + 2, 16, device=dev)
    with pytest.raises(AssertionError):
        non_max_suppression(p, conf_thres=1.5)
    with pytest.raises(AssertionError):
        non_max_suppression(p, iou_thres=-0.1)
    out = non_max_suppression((p, None), nc=3)   # tuple input accepted (utils/ops.py:333-334)
    assert len(out) == 1 and out[0].shape == (0, 8)
    # empty batch / no anchors: the reference's `[zeros((0, 6+nm))] * bs` (utils/ops.py:362)
    assert non_max_suppression(torch.zeros(0, 9, 16, device=dev), nc=3) == []
    out = non_max_suppression(torch.zeros(2, 9, 0, device=dev), nc=3)
    assert len(out) == 2 and all(o.shape == (0, 8) for o in out)
    from ycr_b200.head import decode
    from ycr_b200.ops import detect
    empty = [torch.zeros(0, 46, s, s, device=dev) for s in (20, 10, 5)]
    assert decode(empty, [8, 16, 32], 10, 36).shape == (0, 4 + 10 + 108, 525)
    assert detect(empty, [8, 16, 32], 10, 36, 0.25, 0.7) == []


def test_segment_head_forward_contract():
    """Segment.forward outputs: train -> (feats,5,2); eval -> (allpred,(feats,allpred,1))."""
    from ycr_b200.head import Segment
    dev = _dev()
    torch.manual_seed(0)
    head = Segment(nc=10, nm=36, ch=(32, 64, 128)).to(dev)
    head.stride = torch.tensor([8., 16., 32.])
    head.bias_init()
    x = [torch.randn(2, c, s, s, device=dev) for c, s in ((32, 20), (64, 10), (128, 5))]
    head.train()
    feats, a, b = head([t.clone() for t in x])
    assert (a, b) == (5, 2) and [tuple(f.shape) for f in feats] == [(2, 46, 20, 20), (2, 46, 10, 10), (2, 46, 5, 5)]
    head.eval()
    with torch.no_grad():
        allpred, (feats2, allpred2, one) = head([t.clone() for t in x])
    assert allpred.shape == (2, 4 + 10 + 108, 525) and one == 1 and allpred2 is allpred
    ref = po.decode([f.cpu() for f in feats2], (8, 16, 32), 10, 36)
    assert bool(((allpred.cpu() - ref).abs() <= 1e-5 * ref.abs() + 1e-5).all())


def test_nms_padded_and_compact_row_layouts_agree():
    """The C ABI offers both output layouts (ycr_nms_cfg_t.compact_rows); the Python mirror uses the compact
    one.  Same kept rows, same counts, including images that keep nothing."""
    import ctypes as C
    from ycr_b200 import _lib as L
    from ycr_b200.ops import non_max_suppression
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(9)
    B, nc, R, A = 5, 3, 36, 600
    CH = 4 + nc + 3 * R
    pred = torch.zeros(B, CH, A)
    xy = torch.rand(B, 2, A, generator=g) * 500
    wh = torch.rand(B, 2, A, generator=g) * 80 + 5
    pred[:, 0:2] = xy
    pred[:, 2:4] = xy + wh
    pred[:, 4:4 + nc] = torch.rand(B, nc, A, generator=g) * 0.6
    pred[2, 4:4 + nc] = 0.01                                   # image 2 keeps nothing
    pred[:, 4 + nc:] = torch.rand(B, 3 * R, A, generator=g)
    pred = pred.to(dev)
    ref = non_max_suppression(pred, 0.25, 0.5, nc=nc, max_det=50)
    assert ref[2].shape[0] == 0 and sum(r.shape[0] for r in ref) > 0
    cfg = L.NmsCfg()
    cfg.conf_thres, cfg.iou_thres, cfg.agnostic, cfg.multi_label = 0.25, 0.5, 0, 0
    cfg.max_det, cfg.nc, cfg.max_nms, cfg.max_wh, cfg.classes, cfg.n_classes = 50, nc, 30000, 7680.0, None, 0
    cfg.compact_rows = 0
    lib = L.lib()
    ws = torch.empty(lib.ycr_nms_workspace_bytes(B, A, CH, C.byref(cfg)), dtype=torch.uint8, device=dev)
    rows = torch.full((B, 50, 6 + 3 * R), -1.0, device=dev)
    cnt = torch.empty(B, dtype=torch.int32, device=dev)
    rc = lib.ycr_nms(pred.data_ptr(), B, CH, A, C.byref(cfg), rows.data_ptr(), cnt.data_ptr(), ws.data_ptr(),
                     ws.numel(), L.stream_ptr(dev))
    L.check(rc, "ycr_nms")
    n = cnt.tolist()
    assert n == [r.shape[0] for r in ref]
    for b in range(B):
        assert torch.equal(rows[b, :n[b]], ref[b])
        assert bool((rows[b, n[b]:] == -1.0).all())            # nothing written beyond the kept rows


@pytest.mark.parametrize("imgsz", [384, 320])   # 320: a 10-wide level, the scalar decode provides no by-product
def test_nms_uses_the_decode_by_product_only_when_valid(imgsz):
    """head.decode leaves the per-anchor best class on its output tensor; non_max_suppression starts from it
    when it is handed that very tensor, unmodified - and must not when the tensor was changed in place or is
    another object.  All three routes have to agree with a plain scan of the class rows."""
    from ycr_b200 import synth
    from ycr_b200.head import decode
    from ycr_b200.ops import non_max_suppression
    dev = _dev()
    cfg = synth.PathConfig("h", 4, 0, imgsz, rays=36, nc=12)
    feats = [f.to(dev) for f in synth.make_feats(cfg, 77)]
    out = decode(feats, cfg.strides, cfg.nc, cfg.rays)
    assert getattr(out, "_ycr_best_class", None) is not None
    best, ver, nc = out._ycr_best_class
    conf, j = out[:, 4:4 + cfg.nc].max(1)                       # what utils/ops.py:386 computes
    if imgsz == 384:
        assert torch.equal(best[..., 0].view(torch.float32), conf) and torch.equal(best[..., 1].long(), j)
    else:
        assert bool((best[..., 1] == -1).all())                   # "not provided": NMS scans the class rows
    with_hint = non_max_suppression(out, 0.25, 0.6, nc=cfg.nc)
    plain = non_max_suppression(out.clone(), 0.25, 0.6, nc=cfg.nc)          # a copy carries no by-product
    assert sum(d.shape[0] for d in plain) > 0
    for a, b in zip(with_hint, plain):
        assert torch.equal(a, b)
    out[:, 4:4 + cfg.nc] *= 0.5                                   # in place: the by-product is stale now
    stale = non_max_suppression(out, 0.25, 0.6, nc=cfg.nc)
    fresh = non_max_suppression(out.clone(), 0.25, 0.6, nc=cfg.nc)
    assert [d.shape[0] for d in stale] == [d.shape[0] for d in fresh]
    for a, b in zip(stale, fresh):
        assert torch.equal(a, b)
    assert [d.shape[0] for d in fresh] != [d.shape[0] for d in plain]    # (the change did matter)


@pytest.mark.parametrize("dt", [torch.float32, torch.float16], ids=["fp32", "fp16"])
def test_detect_one_call_equals_decode_plus_nms(dt):
    """Deployment path (ycr_detect): feature maps -> kept rows without the prediction tensor; the rows are
    those of Segment decode + non_max_suppression, bit for bit, also with a class filter and agnostic NMS."""
    from ycr_b200 import synth
    from ycr_b200.head import decode
    from ycr_b200.ops import non_max_suppression, detect
    dev = _dev()
    for (B, S, nc, seed) in ((4, 320, 10, 3), (16, 640, 80, 4)):
        cfg = synth.PathConfig("det", B, 0, S, nc=nc)
        feats = [f.to(dev).to(dt) for f in synth.make_feats(cfg, seed)]
        for kw in (dict(conf_thres=0.25, iou_thres=0.7), dict(conf_thres=0.1, iou_thres=0.45, classes=[1, 3, 5]),
                   dict(conf_thres=0.3, iou_thres=0.6, agnostic=True, max_det=50)):
            want = non_max_suppression(decode(feats, cfg.strides, nc, cfg.rays), nc=nc, **kw)
            got = detect(feats, cfg.strides, nc, cfg.rays, **kw)
            assert [g.shape for g in got] == [w.shape for w in want]
            assert sum(w.shape[0] for w in want) > 0
            for g, w in zip(got, want):
                assert torch.equal(g, w)
