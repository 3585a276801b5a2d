import sys, time; sys.path.insert(0,'/root/repo')
import torch, dataclasses
import ycr_b200
from ycr_b200 import synth
from ycr_b200.head import decode
from ycr_b200.ops import non_max_suppression
dev=torch.device('cuda:0')
cfg=synth.CONFIGS['C3']
for B in (32, 256):
    small = synth.make_feats(synth.PathConfig("gi", 16, 0, cfg.imgsz, rays=36, nc=80), 1001)
    feats=[f.repeat(B//16,1,1,1).to(dev) for f in small]
    pred=decode(feats, cfg.strides, 80, 36)
    for name,kw in (("validator conf .001 multi_label iou .6", dict(conf_thres=0.001, iou_thres=0.6, multi_label=True, max_det=300)),
                    ("predictor conf .25 iou .7", dict(conf_thres=0.25, iou_thres=0.7, max_det=300))):
        for _ in range(3): d=non_max_suppression(pred, nc=80, **kw)
        torch.cuda.synchronize(); t=time.perf_counter()
        for _ in range(10): d=non_max_suppression(pred, nc=80, **kw)
        torch.cuda.synchronize(); dt=(time.perf_counter()-t)/10
        print(f"B={B} {name}: {dt*1e3:.3f} ms per batch, {B/dt:.0f} images/s, kept {sum(x.shape[0] for x in d)/B:.0f}/img")
