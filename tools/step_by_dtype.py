import sys; sys.path.insert(0,'/root/repo')
import torch
import ycr_b200
from ycr_b200 import synth
from ycr_b200.loss import v8SegmentationLoss
import bench
dev=torch.device('cuda:0')
cfg=synth.CONFIGS['C2']
batch, feats = bench.bench_inputs(cfg, 1000)
crit=v8SegmentationLoss(nc=80,nm=36,strides=cfg.strides,device=dev)
for dt in (torch.float16, torch.float32):
    fd=[f.to(dev).to(dt).requires_grad_(True) for f in feats]
    for _ in range(4):
        for f in fd: f.grad=None
        t,i=crit((fd,5,2),batch); t.backward()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        for f in fd: f.grad=None
        t,i=crit((fd,5,2),batch); t.backward()
    e1.record(); torch.cuda.synchronize()
    print(dt, e0.elapsed_time(e1)/20)
