"""Run the C2 step many times on the same inputs and report every distinct (loss, items, tss) that comes out."""
import sys, os; sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import collections
import torch
import ycr_b200
from ycr_b200 import synth
from ycr_b200.loss import v8SegmentationLoss, _SegLossFn
from ycr_b200.tal import gt_struct
mode = sys.argv[1] if len(sys.argv) > 1 else "full"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 150
dev = torch.device('cuda:0')
cfg = synth.CONFIGS['C2']
batch = synth.make_gts(cfg, 305)
feats = [f.to(dev) for f in synth.make_feats_near_gt(cfg, 305, batch)]
crit = v8SegmentationLoss(nc=cfg.nc, nm=cfg.rays, strides=cfg.strides, device=dev)
seen = collections.Counter()
gsum = collections.Counter()
if mode == "packed":
    crit._shapes = [tuple(f.shape[2:]) for f in feats]
    packed, cap = crit.pack_targets(batch, cfg.batch, (640, 640))
outs = []
from ycr_b200 import _lib as L
for it in range(n):
    fl = [f.clone().requires_grad_(True) for f in feats]
    if mode.startswith("poison") and it > 0:   # scratch memory must not carry anything from call to call
        buf = L.Workspace.get("seg_loss", 1, dev)
        if mode == "poison_ff":
            buf.fill_(255)
        elif mode == "poison_00":
            buf.zero_()
        else:
            buf.random_(0, 256)
    if mode == "packed":
        gl, gb, gc = packed.split((1, 4, 720), 2)
        gt, keep = gt_struct(gl, gb, gc, None)
        total, out = _SegLossFn.apply(crit, gt, cap, *fl)
    else:
        crit._shapes = [tuple(f.shape[2:]) for f in fl]
        p, cap2 = crit.pack_targets(batch, cfg.batch, (640, 640))
        gl, gb, gc = p.split((1, 4, 720), 2)
        gt, keep = gt_struct(gl, gb, gc, None)
        total, out = _SegLossFn.apply(crit, gt, cap2, *fl)
        if mode == "full_checkpack":
            outs.append(p.clone())
    if mode != "nobwd":
        total.backward()
    outs.append((out.clone(), [f.grad.double().sum() for f in fl] if mode != "nobwd" else None))
torch.cuda.synchronize()
for o in outs:
    if isinstance(o, tuple):
        seen[tuple(o[0].tolist())] += 1
        if o[1] is not None:
            gsum[tuple(float(x) for x in o[1])] += 1
print(mode, "distinct loss_out:", len(seen), "distinct grad sums:", len(gsum))
for k, v in seen.most_common(6):
    print("  ", v, k)
if mode == "full_checkpack":
    ps = [o for o in outs if not isinstance(o, tuple)]
    print("packed tensors identical:", all(torch.equal(ps[0], q) for q in ps))
