import sys; sys.path.insert(0,".")
import torch
import ycr_b200
from ycr_b200 import synth
from ycr_b200.head import decode
from ycr_b200.ops import non_max_suppression
dev=torch.device("cuda:0")
cfg=synth.CONFIGS["C3"]
small=synth.make_feats(synth.PathConfig("gi",16,0,cfg.imgsz,rays=36,nc=80),1001)
feats=[f.repeat(16,1,1,1).to(dev) for f in small]
for _ in range(3):
    d=non_max_suppression(decode(feats,cfg.strides,80,36),0.25,0.7,nc=80)
torch.cuda.synchronize()
print(sum(x.shape[0] for x in d))
