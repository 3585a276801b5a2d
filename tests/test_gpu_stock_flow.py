"""GPU: the reference's own train / val / predict flow on the rebound kernels (VERDICT r1 'next' item 5).

baseline/_ref is the unmodified reference installed by baseline/install_reference.py (it travels to the GPU box with
the repo).  The test builds the reference's SegmentationModel from its own yolov8-seg.yaml twice - once untouched on
the CPU, once after `ycr_b200.install.install()` on the GPU with the same weights - and runs BASELINE configs[0]
(2 x 3 x 640 x 640, 8 GTs per image): one training step (nn/tasks.py:82,222), one validator batch
(models/yolo/segment/val.py:46-61,149-219) and one predictor postprocess (models/yolo/segment/predict.py:16-44)."""
import sys

import numpy as np
import pytest
import torch

from util import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref():
    from baseline import refload
    if not refload.available():
        pytest.skip("baseline/_ref not installed (python baseline/install_reference.py in the build container)")
    return refload.load()


def test_train_val_predict_flow_runs_on_rebound_kernels(ref):
    import flow_common as fc
    from ycr_b200 import install
    dev = torch.device("cuda:0")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.deterministic = True
    cfg, batch = fc.c1_batch(nc=10, seed=1)

    # ---- the untouched reference on the CPU ----
    torch.manual_seed(0)
    m_ref = fc.build_model(nc=10)
    assert type(m_ref.model[-1]).__module__ == "ultralytics.nn.modules.head"
    state = {k: v.clone() for k, v in m_ref.state_dict().items()}
    m_ref.eval()
    with torch.no_grad():
        allpred_ref = m_ref(batch["img"])[0]        # (before any training forward moves the BatchNorm statistics)
    m_ref.train()
    loss_ref, items_ref = m_ref(batch)
    loss_ref.backward()

    # ---- the same flow after install() ----
    saved = {(m, a): getattr(sys.modules[m], a) for m, a, _, _ in install.PATCHES
             if m in sys.modules and hasattr(sys.modules[m], a)}
    try:
        done = install.install(strict=True)
        assert ("ultralytics.nn.tasks", "Segment") in done and ("ultralytics.utils.ops", "process_mask") in done
        m = fc.build_model(nc=10)
        assert type(m.model[-1]).__module__.startswith("ycr_b200")              # parse_model picked the rebound head
        m.load_state_dict(state, strict=True)                                   # same parameter names and shapes
        m.to(dev).train()
        gbatch = dict(batch)
        gbatch["img"] = batch["img"].to(dev)
        loss, items = m(gbatch)                                                 # BaseModel.forward -> loss -> criterion
        assert type(m.criterion).__module__.startswith("ycr_b200")
        loss.backward()
        # (1) the head outputs of the two builds agree (cuDNN here, MKL there, fp32 both) ...
        with torch.no_grad():
            feats = m._predict_once(gbatch["img"])[0]
            m_ref.train()
            feats_ref = m_ref._predict_once(batch["img"])[0]
        for f, fr in zip(feats, feats_ref):
            assert float((f.cpu() - fr).abs().max()) < 2e-3
        # (2) ... and on the SAME head outputs the rebound criterion returns what the reference's criterion returns
        # (checked below, once the reference's symbols are restored).
        fg = [f.detach().clone().requires_grad_(True) for f in feats]
        loss_g, items_g = m.criterion((fg, 5, 2), gbatch)
        loss_g.backward()
        same_feats = {"feats": [f.detach().cpu() for f in feats], "items": items_g.cpu(), "loss": loss_g.detach().cpu(),
                      "grads": [f.grad.cpu() for f in fg]}
        assert abs(float(loss) - float(loss_ref)) < 0.1 * float(loss_ref)        # same ball park end to end
        assert sum(int(p.grad is not None and bool(torch.isfinite(p.grad).all())) for p in m.parameters()) > 100

        # ---- validator batch ----
        m.load_state_dict(state, strict=True)       # the BatchNorm statistics of the untrained model again
        m.eval()
        with torch.no_grad():
            preds = m(gbatch["img"])
        allpred = preds[0]
        assert allpred.shape == (2, 4 + 10 + 108, 8400) and allpred.shape == allpred_ref.shape
        assert float((allpred.cpu() - allpred_ref).abs().max()) < 2e-3
        v = fc.make_validator(m, dev)
        out = v.postprocess(preds)                                              # rebound non_max_suppression
        assert len(out) == 2 and all(o.shape[1] == 6 + 108 and o.is_cuda for o in out)
        vb = dict(gbatch)
        vb["masks"] = batch["masks"].to(dev).float()
        vb["batch_idx"] = batch["batch_idx"].to(dev)
        vb["cls"] = batch["cls"].to(dev)
        vb["bboxes"] = batch["bboxes"].to(dev)
        v.update_metrics(out, vb)                                               # rebound process_mask + mask_iou
        assert len(v.stats) == 2
        cb, cm, conf, pcls, tcls = v.stats[0]
        assert cb.shape == (out[0].shape[0], 10) and cm.shape == cb.shape and cm.dtype == torch.bool
        # the masks the validator saw are real fills now, not the reference's zeros
        pm = v.process(1, out[0][:, 6:], out[0][:, :4], shape=(640, 640))
        assert pm.shape == (out[0].shape[0], 640, 640) and int(pm.sum()) > 0

        # ---- predictor postprocess ----
        p = fc.make_predictor(m, dev)
        p.args.conf = 0.001
        res = p.postprocess(preds, gbatch["img"], [np.zeros((640, 640, 3), np.uint8)] * 2)
        assert len(res) == 2 and res[0].boxes.data.shape[1] == 6
        assert res[0].masks is not None and res[0].masks.data.shape[1:] == (640, 640)
    finally:
        for (mod, a), val in saved.items():
            setattr(sys.modules[mod], a, val)

    # ---- the reference's own criterion (symbols restored) on the head outputs the GPU model produced ----
    # End to end the two losses differ by a few per cent at most: a freshly initialised head scores all anchors almost
    # alike, so convolution noise of 1e-6 can reorder a per-GT top-10 - the reference's sensitivity, not the kernels'.
    crit_ref = m_ref.init_criterion()
    assert type(crit_ref).__module__ == "ultralytics.utils.loss"
    fl = [f.clone().requires_grad_(True) for f in same_feats["feats"]]
    loss_r, items_r = crit_ref((fl, 5, 2), batch)
    loss_r.backward()
    assert rel_err(same_feats["items"], items_r) < 1e-4, (same_feats["items"], items_r)
    assert rel_err(same_feats["loss"], loss_r.detach()) < 1e-4
    for a, b in zip(same_feats["grads"], fl):
        assert float((a - b.grad).abs().max()) <= 1e-4 * float(b.grad.abs().max())
