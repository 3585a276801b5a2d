import sys, time
sys.path.insert(0, '/root/repo')
import torch
import ycr_b200
from ycr_b200 import synth
from ycr_b200.loss import v8SegmentationLoss
import bench
dev = torch.device('cuda:0')
cfg = synth.CONFIGS['C2']
batch, feats = bench.bench_inputs(cfg, 1000)
fd = [f.to(dev).requires_grad_(True) for f in feats]
crit = v8SegmentationLoss(nc=80, nm=36, strides=cfg.strides, device=dev)
def step():
    for f in fd: f.grad = None
    total, items = crit((fd, 5, 2), batch); total.backward()
for _ in range(5): step()
torch.cuda.synchronize()
# piecewise host timing (sync after each piece to isolate host cost from queueing)
import ctypes as C
T = {}
def tm(name, f, n=50):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): r = f()
    T[name] = (time.perf_counter() - t0) / n * 1e3
    torch.cuda.synchronize()
    return r
crit._shapes = [tuple(f.shape[2:]) for f in fd]
tm('pack_targets(host issue)', lambda: crit.pack_targets(batch, 64, (640, 640)))
packed, cap = crit.pack_targets(batch, 64, (640, 640))
def fwd():
    for f in fd: f.grad = None
    return crit.call_packed(fd, packed, cap)
tm('call_packed fwd (host issue)', fwd)
def fb():
    for f in fd: f.grad = None
    t, i = crit.call_packed(fd, packed, cap); t.backward()
tm('call_packed fwd+bwd (host issue)', fb)
tm('full step (host issue)', step)
segs = batch['segments']
N = batch['batch_idx'].numel()
pin = torch.empty(N * 720).pin_memory()
tm('cat segs into pinned', lambda: torch.cat([t.reshape(-1, 720) for t in segs], 0, out=pin.view(N, 720)))
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(50): step()
torch.cuda.synchronize(); T['full step wall (synced at end)'] = (time.perf_counter() - t0) / 50 * 1e3
for k, v in T.items(): print(f'{k:40s} {v:.3f} ms')
