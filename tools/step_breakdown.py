"""Where the step time goes outside the kernels: the bare C call in a loop (fixed buffers, no autograd), the packed
call through autograd with and without backward, the full call."""
import sys, time; sys.path.insert(0, '/root/repo')
import ctypes as C
import torch
import ycr_b200
from ycr_b200 import synth, _lib as L
from ycr_b200.loss import v8SegmentationLoss
from ycr_b200.tal import gt_struct
import bench
dev = torch.device('cuda:0')
cfg = synth.CONFIGS['C2']
batch, feats = bench.bench_inputs(cfg, 1000)
crit = v8SegmentationLoss(nc=80, nm=36, strides=cfg.strides, device=dev)
fd = [f.to(dev).requires_grad_(True) for f in feats]
lib = L.lib()

def timeit(f, n=30):
    for _ in range(5): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

def full():
    for f in fd: f.grad = None
    t, i = crit((fd, 5, 2), batch); t.backward()
crit._shapes = [tuple(f.shape[2:]) for f in fd]
packed, cap = crit.pack_targets(batch, 64, (640, 640))
def packed_fb():
    for f in fd: f.grad = None
    t, i = crit.call_packed(fd, packed, cap); t.backward()
def packed_f():
    t, i = crit.call_packed(fd, packed, cap)
# bare C call
shapes = [tuple(f.shape[2:]) for f in fd]
cgrid = L.make_grid(shapes, crit.stride_list)
gl, gb, gc = packed.split((1, 4, 720), 2)
gt, keep = gt_struct(gl, gb, gc, None)
grads = [torch.empty_like(f) for f in fd]
loss_out = torch.empty(4, device=dev)
nbytes = lib.ycr_seg_loss_workspace_bytes(C.byref(cgrid), 64, gt.G, C.byref(crit.acfg), cap)
ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
fp, gp = L.ptr_array(fd), L.ptr_array(grads)
st = L.stream_ptr(dev)
def bare():
    rc = lib.ycr_seg_loss_fwd_bwd_dt(C.byref(cgrid), fp, gp, 0, C.byref(gt), C.byref(crit.acfg), C.byref(crit.lcfg),
                                     loss_out.data_ptr(), ws.data_ptr(), ws.numel(), cap, st)
    assert rc == 0
for name, f in (("bare C call", bare), ("call_packed fwd only", packed_f), ("call_packed fwd+bwd", packed_fb), ("full __call__ + backward", full),
                ("bare C call", bare), ("full __call__ + backward", full)):
    print(f"{name:28s} {timeit(f):.4f} ms")
