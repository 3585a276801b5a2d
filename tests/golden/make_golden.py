"""Generate golden vectors by RUNNING THE REFERENCE ITSELF (CPU, fp32) on seeded synthetic inputs.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py
Writes tests/golden/*.npz.  Nothing in tests/, bench.py or smoke() reads /root/reference at run time;
they read these files.

Patches applied to the reference so that it runs at all (all probed):
  * matplotlib is not installed -> a stub package is put on sys.path;
  * `Segment` is bound to the wrong parent class in the snapshot (nn/modules/head.py:436) — we do not
    build the model graph here, we call the hot-path symbols directly:
      ultralytics.utils.loss.v8SegmentationLoss.__call__   (utils/loss.py:808)
      ultralytics.utils.tal.TaskAlignedAssigner.forward    (utils/tal.py:1135)  [captured inside]
      ultralytics.nn.modules.head.Segment.distance2mask    (nn/modules/head.py:461)
      ultralytics.utils.ops.non_max_suppression            (utils/ops.py:285)
"""
import os
import sys
import tempfile
import types
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/ultralytics-main"


def import_reference():
    os.environ.setdefault("YOLO_CONFIG_DIR", tempfile.mkdtemp())
    try:
        import matplotlib  # noqa: F401
    except ImportError:
        stub = tempfile.mkdtemp()
        os.makedirs(os.path.join(stub, "matplotlib"))
        with open(os.path.join(stub, "matplotlib", "__init__.py"), "w") as f:
            f.write("def use(*a, **k):\n    pass\ndef rc(*a, **k):\n    pass\n"
                    "class _F:\n    def __getattr__(self, n):\n        return _F()\n"
                    "    def __call__(self, *a, **k):\n        return _F()\n"
                    "font_manager = _F()\nrcParams = {}\n")
        with open(os.path.join(stub, "matplotlib", "pyplot.py"), "w") as f:
            f.write("def __getattr__(n):\n    def f(*a, **k):\n        return None\n    return f\n")
        sys.path.insert(0, stub)
    sys.path.insert(0, REF)
    import ultralytics.utils.loss as rloss
    import ultralytics.utils.tal as rtal
    import ultralytics.utils.ops as rops
    import ultralytics.nn.modules.head as rhead
    return rloss, rtal, rops, rhead


class _Head(torch.nn.Module):
    def __init__(self, nc, R, strides):
        super().__init__()
        self.nc, self.nm, self.no, self.reg_max = nc, R, nc + R, 16
        self.stride = torch.tensor(strides, dtype=torch.float32)
        self.w = torch.nn.Parameter(torch.zeros(1))


class _Model(torch.nn.Module):
    def __init__(self, nc, R, strides):
        super().__init__()
        self.model = torch.nn.ModuleList([_Head(nc, R, strides)])
        self.args = SimpleNamespace(box=7.5, cls=0.5, dfl=1.5, overlap_mask=True)


def run_loss(rloss, cfg, feats, batch):
    """Reference v8SegmentationLoss fwd+bwd; also captures the assigner's 8-tuple."""
    crit = rloss.v8SegmentationLoss(_Model(cfg.nc, cfg.rays, cfg.strides))
    captured = {}
    orig = crit.assigner.forward

    def spy(*a, **k):
        r = orig(*a, **k)
        captured["out"] = r
        captured["in"] = a
        return r
    crit.assigner.forward = spy
    fl = [f.clone().requires_grad_(True) for f in feats]
    total, items = crit((fl, 5, 2), batch)
    total.backward()
    return total.detach(), items, [f.grad for f in fl], captured


def pack_assign(captured, prefix, store_dense):
    tl, tb, ts, mp, tgi, gd, cen, fg = captured["out"]
    d = {
        prefix + "target_gt_idx": tgi.numpy().astype(np.int32),
        prefix + "fg_mask": fg.numpy(),
        prefix + "target_labels": tl.numpy().astype(np.int32),
        prefix + "gt_dist": gd.numpy(),
        prefix + "centerness": cen.numpy(),
        prefix + "mask_pos_nz": torch.nonzero(mp).numpy().astype(np.int32),
        prefix + "target_scores_nz_idx": torch.nonzero(ts).numpy().astype(np.int32),
        prefix + "target_scores_nz_val": ts[ts != 0].numpy(),
        prefix + "target_bboxes_fg": tb[fg].numpy(),
    }
    if store_dense:
        d[prefix + "target_scores"] = ts.numpy()
    return d


def grad_digest(grads, fg_rows=None):
    """Dense grads are as large as the inputs; keep sums, abs-sums and a strided sample."""
    d = {}
    for li, g in enumerate(grads):
        flat = g.flatten().double()
        d[f"grad{li}_sum"] = np.float64(flat.sum())
        d[f"grad{li}_abssum"] = np.float64(flat.abs().sum())
        d[f"grad{li}_sample"] = g.flatten()[::97].numpy()
        d[f"grad{li}_ray_nz_idx"] = torch.nonzero(g[:, :36].flatten()).flatten().numpy().astype(np.int64) \
            if g.shape[1] > 36 else np.zeros(0, np.int64)
        d[f"grad{li}_ray_nz_val"] = g[:, :36].flatten()[g[:, :36].flatten() != 0].numpy()
    return d


def main():
    from importlib import import_module
    sys.path.insert(0, ROOT)
    import ycr_b200  # noqa: F401
    synth = import_module("ycr_b200.synth")
    rloss, rtal, rops, rhead = import_reference()
    torch.set_num_threads(os.cpu_count())
    torch.manual_seed(0)

    # ---------------- training path ----------------
    train_cases = [
        # name, cfg, seed, ragged, near_gt, store full inputs/outputs
        ("train_s160", synth.PathConfig("s160", 2, 4, 160, nc=10), 11, False, False, True),
        ("train_s320_ragged", synth.PathConfig("s320", 3, 6, 320, nc=10), 12, True, True, True),
        ("train_c1", synth.CONFIGS["C1"], 13, False, False, False),
        ("train_c1_near", synth.CONFIGS["C1"], 14, False, True, False),
    ]
    for name, cfg, seed, ragged, near, full in train_cases:
        batch = synth.make_gts(cfg, seed, ragged=ragged)
        feats = synth.make_feats_near_gt(cfg, seed, batch) if near else synth.make_feats(cfg, seed)
        total, items, grads, cap = run_loss(rloss, cfg, feats, batch)
        d = {"seed": seed, "ragged": ragged, "near": near,
             "cfg": np.array([cfg.batch, cfg.gts, cfg.imgsz, cfg.rays, cfg.nc]),
             "loss": total.numpy(), "loss_items": items.numpy()}
        d.update(pack_assign(cap, "asg_", full))
        d.update(grad_digest(grads))
        if full:
            for li, (f, g) in enumerate(zip(feats, grads)):
                d[f"feat{li}"] = f.numpy()
                d[f"grad{li}"] = g.numpy()
            d["batch_idx"] = batch["batch_idx"].numpy()
            d["cls"] = batch["cls"].numpy()
            d["bboxes"] = batch["bboxes"].numpy()
            d["segments"] = torch.cat(batch["segments"]).numpy()
        else:
            d["feat_digest"] = np.array([float(f.double().sum()) for f in feats])
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
        print(name, "loss", float(total), "items", items.tolist(),
              "P", int(cap["out"][5].shape[0]), flush=True)

    # ---------------- known-answer: circle GT centred on an anchor ----------------
    asg = rtal.TaskAlignedAssigner(topk=10, num_classes=3, alpha=0.5, beta=4.0)
    level_shapes = [(80, 80), (40, 40), (20, 20)]
    anc, st, ss = rtal.make_anchors_polar([torch.zeros(1, 1, h, w) for h, w in level_shapes], [8, 16, 32])
    A = anc.shape[0]
    t = np.linspace(0, 2 * np.pi, 360, endpoint=False)
    circ = np.stack([324 + 50 * np.cos(t), 324 + 50 * np.sin(t)], 1).astype(np.float32)
    circ = synth.resample_closed(circ)
    gt_coor = torch.from_numpy(circ).view(1, 1, 720)
    gt_boxes = torch.tensor([[[274., 274., 374., 374.]]])
    gt_labels = torch.tensor([[[1.]]])
    mask_gt = torch.ones(1, 1, 1)
    g = torch.Generator().manual_seed(5)
    pd_scores = torch.rand(1, A, 3, generator=g) * 0.5 + 0.1
    pd_rays = torch.full((1, A, 36), 50.0) + torch.rand(1, A, 36, generator=g) * 20
    ci = int(((anc * st - torch.tensor([324., 324.])).abs().sum(1)).argmin())
    pd_rays[0, ci] = 50.0
    pd_scores[0, ci, 1] = 0.95
    out = asg(pd_scores, pd_rays, anc * st, gt_labels, gt_boxes, mask_gt, gt_coor, st, ss, 0,
              torch.tensor([640., 640.]))
    np.savez_compressed(os.path.join(HERE, "kat_circle.npz"),
                        pd_scores=pd_scores.numpy(), pd_rays=pd_rays.numpy(), gt_coor=gt_coor.numpy(),
                        centre_anchor=ci, **pack_assign({"out": out}, "asg_", True))
    print("kat_circle: P", out[5].shape[0], "centre score", float(out[2][0, ci, 1]),
          "gt_dist range", float(out[5].min()), float(out[5].max()), flush=True)

    # ---------------- dormant box terms: BboxLoss (CIoU + DFL), utils/loss.py:53-87 ----------------
    g = torch.Generator().manual_seed(123)
    B, A, nc, reg = 2, 525, 10, 15
    from oracle import polar_oracle as po
    anc, _ = po.make_anchors([(20, 20), (10, 10), (5, 5)], [8, 16, 32])   # grid units, as the reference passes them
    xy = torch.rand(B, A, 2, generator=g) * 12 + 2
    wh = torch.rand(B, A, 2, generator=g) * 6 + 1
    tb = torch.cat([xy - wh, xy + wh], -1)
    pb = (tb + torch.randn(B, A, 4, generator=g) * 0.7).requires_grad_(True)
    pd = torch.randn(B, A, 4 * (reg + 1), generator=g).requires_grad_(True)
    fg = torch.rand(B, A, generator=g) < 0.08
    ts = torch.zeros(B, A, nc)
    ts[fg] = torch.rand(int(fg.sum()), nc, generator=g) * (torch.rand(int(fg.sum()), nc, generator=g) < 0.15)
    tss = max(ts.sum(), 1)
    li, ld = rloss.BboxLoss(reg, use_dfl=True)(pd, pb, anc, tb, ts, tss, fg)
    (li * 7.5 + ld * 1.5).backward()
    np.savez_compressed(os.path.join(HERE, "bbox_loss.npz"), pred_dist=pd.detach().numpy(),
                        pred_bboxes=pb.detach().numpy(), anchor_points=anc.numpy(), target_bboxes=tb.numpy(),
                        target_scores=ts.numpy(), fg_mask=fg.numpy(), tss=np.float32(tss),
                        loss_iou=li.detach().numpy(), loss_dfl=ld.detach().numpy(), grad_dist=pd.grad.numpy(),
                        grad_bboxes=pb.grad.numpy(), gains=np.array([7.5, 1.5], np.float32))
    print("bbox_loss", float(li), float(ld), int(fg.sum()), flush=True)

    # ---------------- contour resampling (utils/ops.py:676-693), ragged polygons ----------------
    rng = np.random.default_rng(7)
    polys = [rng.uniform(0, 1, size=(int(m), 2)).astype(np.float32) for m in (3, 4, 7, 16, 33, 100, 359, 360, 361, 500)]
    res = rops.resample_segments([p.copy() for p in polys], n=360)
    np.savez_compressed(os.path.join(HERE, "resample.npz"), sizes=np.array([len(p) for p in polys]),
                        points=np.concatenate(polys), out=np.stack(res))
    print("resample", len(polys), res[0].dtype, flush=True)

    # ---------------- inference path ----------------
    infer_cases = [
        ("infer_s160", synth.PathConfig("s160", 2, 0, 160, nc=10), 21, True),
        ("infer_s320", synth.PathConfig("s320", 4, 0, 320, nc=10), 22, False),
        ("infer_c3small", synth.PathConfig("c3s", 4, 0, 640, nc=80), 23, False),
    ]
    for name, cfg, seed, full in infer_cases:
        feats = synth.make_feats(cfg, seed)
        head = SimpleNamespace(no=cfg.rays + cfg.nc, nm=cfg.rays, nc=cfg.nc)
        anc, st = rhead.Segment.make_anchors(head, feats, cfg.strides)
        allpred = rhead.Segment.distance2mask(head, anc * st, [f.clone() for f in feats], st)
        d = {"seed": seed, "cfg": np.array([cfg.batch, cfg.gts, cfg.imgsz, cfg.rays, cfg.nc]),
             "allpred_sum": np.float64(allpred.double().sum()),
             "allpred_sample": allpred.flatten()[::101].numpy()}
        if full:
            d["allpred"] = allpred.numpy()
            for li, f in enumerate(feats):
                d[f"feat{li}"] = f.numpy()
        for tag, kw in (("best", dict(conf_thres=0.25, iou_thres=0.7, multi_label=False)),
                        ("multi", dict(conf_thres=0.25, iou_thres=0.7, multi_label=True)),
                        ("agn", dict(conf_thres=0.4, iou_thres=0.5, agnostic=True, max_det=50))):
            dets = rops.non_max_suppression(allpred, nc=cfg.nc, **kw)
            d[f"nms_{tag}_counts"] = np.array([x.shape[0] for x in dets])
            d[f"nms_{tag}_rows"] = torch.cat(dets).numpy() if len(dets) else np.zeros((0, 6))
            print(name, tag, [x.shape[0] for x in dets], flush=True)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)


if __name__ == "__main__":
    main()
