import sys; sys.path.insert(0,'/root/repo')
import torch, ctypes as C
import ycr_b200
from ycr_b200 import synth, _lib as L
from ycr_b200.loss import v8SegmentationLoss
import bench
dev=torch.device('cuda:0')
cfg=synth.CONFIGS['C2']
batch, feats = bench.bench_inputs(cfg, 1000)
crit=v8SegmentationLoss(nc=80,nm=36,strides=cfg.strides,device=dev)
lib=L.lib()
names=["gt_setup","cand_overlaps","topk","resolve","positives","loss_stream","finalize"]
for dt in (torch.float32, torch.float16):
    fd=[f.to(dev).to(dt).requires_grad_(True) for f in feats]
    for _ in range(4):
        for f in fd: f.grad=None
        t,i=crit((fd,5,2),batch); t.backward()
    torch.cuda.synchronize()
    lib.ycr_profile_select(0xFFFFFFFF); lib.ycr_profile_begin(10*12+64)
    for _ in range(10):
        for f in fd: f.grad=None
        t,i=crit((fd,5,2),batch); t.backward()
    torch.cuda.synchronize()
    sums=(C.c_float*16)(); counts=(C.c_int*16)()
    lib.ycr_profile_end(sums,counts)
    print(sys.argv[1], dt, "loss_stream %.1f us"%(1e3*sums[5]/counts[5]), "K1 %.1f"%(1e3*sums[1]/counts[1]))
