#!/usr/bin/env python
"""Benchmark of the polar-contour hot path (BASELINE.json metric: assign+polar-loss images/sec @640).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C2]

A "step" is one pass of v8SegmentationLoss (assignment + polar targets + loss, forward AND the gradient
w.r.t. the head outputs) over one batch of synthetic head outputs and GTs.  Workload = BASELINE.json
configs[1]: batch 64 @640, 20 GTs/img, 36 rays, nc=80 per GPU (weak scaling: every rank owns a full
batch, as the reference's DDP does; the path has no data-path collective).

  value   images/s with inputs resident in HBM, CUDA events, max over ranks
  e2e     images/s through the public API with HOST inputs: pinned feature maps and the CPU batch dict
          are copied H2D inside the timed region and the loss is read back D2H every step
  roofline / kernels   per-kernel CUDA-event times from the library's own hooks, taken in the timed region
  infer   decode + NMS images/s on config C3 (batch 256 @640, conf .25 / IoU .7), same method
  cpu_baseline   the oracle port (torch-CPU restatement of the reference) on a bounded sample

`--impl reference` times the reference's own v8SegmentationLoss (installed into the git-ignored baseline/_ref by
baseline/install_reference.py, so it travels to the GPU box) on the host cores, on the first images of the same
batch; the oracle port stands in when baseline/_ref is absent or the workload has 72 rays."""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled every ~2 ms through NVML (nvidia-ml-py) by a host thread while
    the timed region runs; the same fields `nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,
    clocks_event_reasons.*` prints."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.sm, self.bits = [], 0
        self.h = None
        self.stop_flag = False
        self.thread = None
        self.sm_max = None
        self.period = float(os.environ.get("YCR_CLOCK_PERIOD_MS", "2")) * 1e-3
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def _loop(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.bits |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception:
                pass
            time.sleep(self.period)

    def start(self):
        if self.h is None:
            return
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def stop(self):
        if self.h is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + getattr(self, "err", "")]}
        self.stop_flag = True
        self.thread.join(timeout=1.0)
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.sm_max,
                "reasons": sorted(n for bit, n in self.REASONS.items() if self.bits & bit), "samples": len(sm)}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference itself (baseline/_ref, installed by baseline/install_reference.py) when it travelled with
# the repo, else the oracle port; bounded sample of the same workload, same input generators as the GPU arm
# ------------------------------------------------------------------------------------------------
K1_EVERY = 4   # the dominant kernel carries CUDA events on every K1_EVERY-th step of the timed region


def bench_inputs(cfg, seed, n_images=None):
    """Synthetic inputs of one rank: the batch dict and the head feature maps (random head outputs,
    16 distinct images tiled to the batch - both arms use this generator)."""
    from ycr_b200 import synth
    B = cfg.batch if n_images is None else n_images
    sub = synth.PathConfig("gen", B, cfg.gts, cfg.imgsz, rays=cfg.rays, nc=cfg.nc)
    batch = synth.make_gts(sub, seed)
    gen_cfg = synth.PathConfig("gen", min(B, 16), cfg.gts, cfg.imgsz, rays=cfg.rays, nc=cfg.nc)
    reps = (B + gen_cfg.batch - 1) // gen_cfg.batch
    small = synth.make_feats(gen_cfg, seed)
    feats = [torch.cat([f.roll(k, 0) for k in range(reps)], 0)[:B].contiguous() for f in small]
    return batch, feats


def cpu_train_sample(cfg, n_images, repeats, seed=1000):
    """-> (times, kind): fwd+bwd of v8SegmentationLoss on the first n_images of the GPU arm's rank-0 batch."""
    from baseline import refload
    batch, feats = bench_inputs(cfg, seed, n_images)
    times = []
    if cfg.rays == 36 and refload.available():   # (the reference hard-codes 36 rays, utils/tal.py:1178)
        crit = refload.reference_criterion(cfg.nc, cfg.rays, cfg.strides)
        for _ in range(repeats):
            fl = [f.clone().requires_grad_(True) for f in feats]
            t0 = time.perf_counter()
            total, items = crit((fl, 5, 2), batch)
            total.backward()
            times.append(time.perf_counter() - t0)
        return times, "reference"
    from oracle import polar_oracle as po
    for _ in range(repeats):
        t0 = time.perf_counter()
        po.seg_loss(feats, batch, cfg.strides, cfg.nc, cfg.rays, with_grad=True)
        times.append(time.perf_counter() - t0)
    return times, "port"


def run_reference(args):
    rank, world, local = dist_env()
    if rank != 0:
        return
    from ycr_b200 import synth
    cfg = synth.CONFIGS[args.workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n_img = args.cpu_images
    times, kind = cpu_train_sample(cfg, n_img, args.warmup + args.steps)
    timed = times[args.warmup:]
    ms = 1e3 * sum(timed) / len(timed)
    val = n_img / (ms / 1e3)
    what = ("the reference's v8SegmentationLoss.__call__ fwd+bwd (baseline/_ref, torch CPU)" if kind == "reference"
            else "oracle/polar_oracle.seg_loss fwd+bwd (port: baseline/_ref absent or rays != 36)")
    line = {
        "impl": "reference", "metric": "assign+polar-loss images/sec @640", "value": val, "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.workload, cfg),
                   "sample": f"{n_img} images per step (bounded sample of the GPU arm's rank-0 batch, same generators)"},
        "cpu_baseline": {"value": val, "unit": "images/s", "cores": cores, "kind": kind,
                         "sample": f"{n_img} images/step x {args.steps} steps, {what}"},
        "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_name(name, cfg):
    return (f"{name}: v8SegmentationLoss.__call__ fwd+bwd from the head feature maps and the batch dict, "
            f"batch {cfg.batch}/GPU @{cfg.imgsz}, {cfg.gts} GTs/img, {cfg.rays} rays, nc={cfg.nc}, A={cfg.anchors}")


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(gpu_index):
    """Pin this process to the CPUs NVML reports as local to its GPU before any pinned host buffer is allocated
    (first touch then places the staging memory on that NUMA node: with one rank per GPU the host-to-device
    copies of the e2e leg do not cross the socket interconnect).  Best effort."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        n_words = (os.cpu_count() + 63) // 64
        masks = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = {64 * w + bit for w, m in enumerate(masks) for bit in range(64) if (int(m) >> bit) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass


def run_dp_step(cfg, dev, world, local, batch, steps):
    """BASELINE configs[4]: the data-parallel training step around the path.  The Segment head of a yolov8x-sized neck
    (cuDNN convolutions under autocast, as the reference trains) feeds the fused loss; with more than one rank the head
    is wrapped in DistributedDataParallel, so the bucketed NCCL all-reduce of its 35 MB of gradients runs during the
    backward pass, `loss *= world_size` as engine/trainer.py:365.  The step is timed with and without the all-reduce
    (`no_sync`): the difference is the part of the collective the backward pass does not hide."""
    import contextlib
    from ycr_b200.head import Segment
    from ycr_b200.loss import v8SegmentationLoss
    B = cfg.batch
    ch = (320, 640, 640)   # neck widths of yolov8x at the three head levels
    torch.manual_seed(1234)
    head = Segment(nc=cfg.nc, nm=cfg.rays, ch=ch).to(dev)
    head.stride = torch.tensor(cfg.strides, dtype=torch.float32)
    head.bias_init()
    head.train()
    model = head
    if world > 1:
        from torch.nn.parallel import DistributedDataParallel as DDP
        model = DDP(head, device_ids=[local])
    crit = v8SegmentationLoss(nc=cfg.nc, nm=cfg.rays, strides=cfg.strides, device=dev)
    g = torch.Generator(device=dev).manual_seed(77 + local)
    x = [torch.randn(B, c, h, w, device=dev, dtype=torch.float16, generator=g) * 0.5
         for c, (h, w) in zip(ch, cfg.level_shapes)]
    n_param = sum(p.numel() for p in head.parameters())

    def step(sync=True):
        model.zero_grad(set_to_none=True)
        ctx = contextlib.nullcontext() if (sync or world == 1) else model.no_sync()
        with ctx:
            with torch.autocast("cuda", dtype=torch.float16):
                feats, _, _ = model(x)
            loss, items = crit((feats, 5, 2), batch)       # fp16 maps, read in place
            if world > 1:
                loss = loss * world
            loss.backward()

    def timed(sync):
        for _ in range(3):
            step(sync)
        torch.cuda.synchronize()
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step(sync)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    ms_sync = timed(True)
    ms_nosync = timed(False) if world > 1 else ms_sync
    del x, model, head
    torch.cuda.empty_cache()
    return ms_sync, ms_nosync, n_param


def run_ours(args):
    rank, world, local = dist_env()
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback in the product)"
    # libraries (NCCL's version banner) write to fd 1; keep stdout clean for the one JSON line
    sys.stdout.flush()
    _real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.dup2(_real_stdout, 1)
        print(json.dumps(obj), flush=True)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    use_dist = world > 1
    if use_dist and not os.environ.get("YCR_NO_BIND"):   # (at N=1 the CPU baseline leg wants every host core)
        bind_to_gpu_numa_node(local)
    if use_dist:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    import ycr_b200  # noqa: F401
    from ycr_b200 import synth, _lib as L
    from ycr_b200.loss import v8SegmentationLoss
    from ycr_b200.head import decode
    from ycr_b200.ops import non_max_suppression
    lib = L.lib()
    cfg = synth.CONFIGS[args.workload]
    B, G, R, nc = cfg.batch, cfg.gts, cfg.rays, cfg.nc
    A = cfg.anchors

    # synthetic inputs: weak scaling keeps the work per GPU fixed, so every rank draws the SAME synthetic batch (the cost
    # of the candidate kernel follows the GT shapes: with seeds 1000+rank the ranks' kernel times spread over 1.02-1.07 ms
    # and the max over ranks measured the heaviest draw against N=1's lightest).  YCR_BENCH_RANK_SEEDS=1 restores that.
    seed = 1000 + (rank if os.environ.get("YCR_BENCH_RANK_SEEDS") else 0)
    batch, feats_cpu = bench_inputs(cfg, seed)
    feats_h = [f.pin_memory() for f in feats_cpu]
    feats_d = [f.to(dev).requires_grad_(True) for f in feats_h]
    crit = v8SegmentationLoss(nc=nc, nm=R, strides=cfg.strides, device=dev)
    in_bytes = sum(f.numel() * 4 for f in feats_h)

    def step_resident():
        # from the head feature maps (resident, requires_grad) + the batch dict (on the CPU,
        # as the dataloader leaves it) to loss.backward() finished - GT packing is inside the region
        for f in feats_d:
            f.grad = None
        total, items = crit((feats_d, 5, 2), batch)
        total.backward()
        return total

    def barrier():
        if use_dist:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    if args.quick_infer:
        icfg = synth.CONFIGS["C3"]
        ismall = synth.make_feats(synth.PathConfig("gi", 16, 0, icfg.imgsz, rays=R, nc=nc), seed + 1)
        ifeats = [torch.cat([f.roll(k, 0) for k in range(icfg.batch // 16)], 0).contiguous().to(dev) for f in ismall]
        L.check(lib.ycr_profile_begin(256), "ycr_profile_begin")
        for _ in range(args.warmup + args.steps):
            non_max_suppression(decode(ifeats, icfg.strides, nc, R), 0.25, 0.7, nc=nc, max_det=300)
        torch.cuda.synchronize()
        sums = (C.c_float * 16)()
        counts = (C.c_int * 16)()
        L.check(lib.ycr_profile_end(sums, counts), "ycr_profile_end")
        names = {7: "decode", 8: "nms_filter", 9: "nms_sort", 10: "nms_suppress"}
        emit({"quick_infer": True, "kernels_ms": {n: sums[i] / counts[i] for i, n in names.items() if counts[i]}})
        return
    for _ in range(max(args.warmup, 3)):
        step_resident()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # the timed region carries the events of the dominant kernel only (quick mode: of every kernel); the
    # per-kernel breakdown of the others comes from a second, untimed pass
    L.check(lib.ycr_profile_select(0xFFFFFFFF if args.quick else (1 << 1)), "ycr_profile_select")
    L.check(lib.ycr_profile_begin(args.steps * 24 + 64), "ycr_profile_begin")
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    t_host0 = time.perf_counter()
    for i in range(args.steps):
        # the two event records around the dominant kernel cost ~19 us of stream time per step (they end the chain of
        # dependent launches on both sides): they are placed around every K1_EVERY-th launch of the timed region
        if not args.quick:
            lib.ycr_profile_select((1 << 1) if i % K1_EVERY == 0 else 0)
        step_resident()
    host_issue_ms = (time.perf_counter() - t_host0) * 1e3 / args.steps   # host time to ISSUE a step (no sync inside)
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    # the same steps with the GT rows already packed on the device (what round 1 timed): shows what the packing costs
    crit._shapes = [tuple(f.shape[2:]) for f in feats_d]
    packed, cap = crit.pack_targets(batch, B, (cfg.imgsz, cfg.imgsz))
    torch.cuda.synchronize()
    pp0, pp1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pp0.record()
    for i in range(args.steps):
        if not args.quick:
            lib.ycr_profile_select((1 << 1) if i % K1_EVERY == 0 else 0)   # as in the timed region
        for f in feats_d:
            f.grad = None
        tot_pp, _ = crit.call_packed(feats_d, packed, cap)
        tot_pp.backward()
    pp1.record()
    torch.cuda.synchronize()
    ms_prepacked = pp0.elapsed_time(pp1) / args.steps
    sums = (C.c_float * 16)()
    counts = (C.c_int * 16)()
    L.check(lib.ycr_profile_end(sums, counts), "ycr_profile_end")
    clocks = sampler.stop() if rank == 0 else None
    L.check(lib.ycr_profile_select(0xFFFFFFFF), "ycr_profile_select")
    k1_samples = counts[1]
    if not args.quick:
        k1_sum, k1_cnt = sums[1], counts[1]
        n_bd = min(args.steps, 10)
        L.check(lib.ycr_profile_begin(n_bd * 12 + 64), "ycr_profile_begin")
        for _ in range(n_bd):
            step_resident()
        torch.cuda.synchronize()
        L.check(lib.ycr_profile_end(sums, counts), "ycr_profile_end")
        sums[1], counts[1] = k1_sum, k1_cnt     # the dominant kernel keeps its timed-region average

    if args.quick:
        names = ["gt_setup", "cand_overlaps", "topk", "resolve", "positives", "loss_stream", "finalize"]
        if rank == 0:
            st = (C.c_ulonglong * 4)()
            lib.ycr_debug_stats(st, 1)
            emit({"quick": True, "ms_per_step": ms_total / args.steps,
                  "stats_per_candidate": {"candidates": st[0], "queued_pairs": st[1] / max(st[0], 1),
                                          "scan_pairs": st[2] / max(st[0], 1)},
                  "kernels_ms": {names[i]: sums[i] / counts[i] for i in range(7) if counts[i]}})
        return
    # ---- e2e: host inputs, H2D inside the timed region, loss read back ----
    gt_rows_bytes = int(batch["batch_idx"].numel()) * 726 * 4

    # Double-buffered host inputs: every step copies its own 253 MB (plus the GT rows) from pinned host memory,
    # but the copy of step k+1 is issued - on a second stream, behind step k's small GT copy - before the host
    # waits for step k's loss, so the copy engine never idles while the kernels run.
    copy_stream = torch.cuda.Stream(dev)
    pending = {}

    def run_e2e(host_feats, n_steps):
        """K steps through the public API from pinned HOST feature maps (copied H2D inside the timed region, double
        buffered) and the CPU batch dict; the loss is read back every step.  -> milliseconds for n_steps."""
        def issue():
            gt_ev = crit.last_gt_copy_event   # this step's GT rows go first (its kernels wait for them)
            if gt_ev is not None:
                copy_stream.wait_event(gt_ev)
            with torch.cuda.stream(copy_stream):
                fd = [f.to(dev, non_blocking=True) for f in host_feats]
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            pending["next"] = (fd, ev)

        def one():
            fd, ev = pending["next"]
            cur = torch.cuda.current_stream(dev)
            cur.wait_event(ev)
            for f in fd:
                f.record_stream(cur)
            fd = [f.requires_grad_(True) for f in fd]
            total, items = crit((fd, 5, 2), batch)
            total.backward()
            issue()
            return float(total.detach())  # D2H read of the step's result

        issue()
        for _ in range(3):
            one()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n_steps):
            one()
        e1.record()
        barrier()
        pending.clear()
        return e0.elapsed_time(e1)

    e_steps = max(3, min(args.steps, 10))
    ms_e2e = run_e2e(feats_h, e_steps)

    # ---- the same with fp16 head outputs, what the head emits under autocast (the reference's default,
    # engine/trainer.py:332): the kernels read the half maps in place and write half gradients ----
    feats_h16 = [f.half().pin_memory() for f in feats_cpu]
    feats_d16 = [f.to(dev).requires_grad_(True) for f in feats_h16]

    def step_resident16():
        for f in feats_d16:
            f.grad = None
        total, items = crit((feats_d16, 5, 2), batch)
        total.backward()

    for _ in range(3):
        step_resident16()
    barrier()
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h_steps = max(3, min(args.steps, 20))
    h0.record()
    for _ in range(h_steps):
        step_resident16()
    h1.record()
    barrier()
    ms_res16 = h0.elapsed_time(h1)
    ms_e2e16 = run_e2e(feats_h16, e_steps)
    del feats_d16

    # ---- inference path (config C3), reported alongside ----
    icfg = synth.CONFIGS["C3"]
    ib = icfg.batch
    ismall = synth.make_feats(synth.PathConfig("gi", 16, 0, icfg.imgsz, rays=R, nc=nc), seed + 1)
    ifeats = [torch.cat([f.roll(k, 0) for k in range(ib // 16)], 0).contiguous().to(dev) for f in ismall]

    def step_infer():
        allpred = decode(ifeats, icfg.strides, nc, R)
        return non_max_suppression(allpred, 0.25, 0.7, nc=nc, max_det=300)

    for _ in range(3):
        step_infer()
    barrier()
    L.check(lib.ycr_profile_begin(64), "ycr_profile_begin")
    i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    i_steps = 5
    i0.record()
    for _ in range(i_steps):
        dets = step_infer()
    i1.record()
    barrier()
    ms_inf = i0.elapsed_time(i1)
    isums = (C.c_float * 16)()
    icounts = (C.c_int * 16)()
    L.check(lib.ycr_profile_end(isums, icounts), "ycr_profile_end")
    kept = sum(d.shape[0] for d in dets) / ib
    # the deployment form: feature maps -> kept rows in one call, no prediction tensor (ops.detect / ycr_detect)
    from ycr_b200.ops import detect
    for _ in range(3):
        detect(ifeats, icfg.strides, nc, R, 0.25, 0.7, max_det=300)
    barrier()
    j0, j1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    j0.record()
    for _ in range(i_steps):
        dets_d = detect(ifeats, icfg.strides, nc, R, 0.25, 0.7, max_det=300)
    j1.record()
    barrier()
    ms_det = j0.elapsed_time(j1)
    assert sum(d.shape[0] for d in dets_d) == sum(d.shape[0] for d in dets)
    # ... and with the packed output of the library call itself (rows back to back + device counts: no host read)
    barrier()
    j0.record()
    for _ in range(i_steps):
        rows_p, counts_p = detect(ifeats, icfg.strides, nc, R, 0.25, 0.7, max_det=300, packed=True)
    j1.record()
    barrier()
    ms_det_packed = j0.elapsed_time(j1)
    assert int(counts_p.sum()) == sum(d.shape[0] for d in dets)

    # ---- data-parallel training step (configs[4]), every rank ----
    dp_steps = max(3, min(args.steps, 10))
    ms_dp, ms_dp_nosync, dp_params = run_dp_step(cfg, dev, world, local, batch, dp_steps)

    # ---- reduce over ranks (max time) ----
    k1_ms_rank = (sums[1] / counts[1]) if counts[1] else 0.0
    t = torch.tensor([ms_total, ms_e2e, ms_inf, ms_res16, ms_e2e16, ms_dp, ms_dp_nosync, ms_prepacked, host_issue_ms,
                      k1_ms_rank], device=dev, dtype=torch.float64)
    per_rank = None
    if use_dist:
        import torch.distributed as dist
        gathered = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(gathered, t)
        per_rank = {"ms_per_step": [float(g[0]) / args.steps for g in gathered],
                    "e2e_ms_per_step": [float(g[1]) / e_steps for g in gathered],
                    "ms_per_step_gt_prepacked": [float(g[7]) for g in gathered],
                    "host_issue_ms_per_step": [float(g[8]) for g in gathered],
                    "cand_overlaps_ms": [float(g[9]) for g in gathered]}
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e, ms_inf, ms_res16, ms_e2e16, ms_dp, ms_dp_nosync = [float(x) for x in t.tolist()[:7]]
    if rank != 0:
        if use_dist:
            import torch.distributed as dist
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peaks()
    # DRAM traffic of the dominant kernel: what an `ncu --set full` capture of this build measured (profiles/),
    # never a number made up in the run; null when no capture of the current round exists
    traffic, traffic_src = {}, None
    try:
        tf = sorted(f for f in os.listdir(os.path.join(ROOT, "profiles")) if f.endswith("_traffic.json"))
        if tf:
            traffic = json.load(open(os.path.join(ROOT, "profiles", tf[-1])))
            traffic_src = "profiles/" + tf[-1]
    except Exception:
        pass
    ms_step = ms_total / args.steps
    value = world * B / (ms_step / 1e3)
    e2e_val = world * B / (ms_e2e / e_steps / 1e3)
    names = ["gt_setup", "cand_overlaps", "topk", "resolve", "positives", "loss_stream", "finalize", "decode",
             "nms_filter", "nms_sort", "nms_suppress"]
    kern = {names[i]: (sums[i] / counts[i]) for i in range(7) if counts[i]}
    ikern = {names[i]: (isums[i] / icounts[i]) for i in range(7, 11) if icounts[i]}
    dom = max(kern, key=kern.get)
    bytes_img = 2 * 4 * A * (R + nc) + 4 * G * (5 + 720)            # SURVEY.md §8-d, per image
    stream_bytes = 2 * 4 * A * (R + nc) * B                          # what loss_stream itself must move
    path_gbs = bytes_img * B / (ms_step * 1e-3) / 1e9
    dom_gbs = bytes_img * B / (kern[dom] * 1e-3) / 1e9
    inf_bytes_img = 4 * A * (R + nc) + 4 * A * (4 + nc + 3 * R) + 4 * kept * (6 + 3 * R)
    inf_val = world * ib / (ms_inf / i_steps / 1e3)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cpu_times, cpu_kind = cpu_train_sample(cfg, args.cpu_images, 3) if world == 1 else (None, None)
    cpu_val = (args.cpu_images / (sum(cpu_times[1:]) / len(cpu_times[1:]))) if cpu_times else None
    line = {
        "metric": "assign+polar-loss images/sec @640", "value": value, "unit": "images/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.workload, cfg),
                   "l2": f"inputs per step {in_bytes / 1e6:.0f} MB + grads of the same size > 126 MB L2",
                   "parallelism": f"dp{world}, no data-path collective",
                   "per_rank_data": ("different seed per rank" if os.environ.get("YCR_BENCH_RANK_SEEDS") else
                                     "the same synthetic batch on every rank (equal work per GPU)")},
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": dom_gbs, "peak": peak, "unit": "GB/s",
                     "frac": dom_gbs / peak, "traffic": traffic.get(dom) if args.workload == "C2" else None,
                     "traffic_source": traffic_src if (args.workload == "C2" and traffic.get(dom)) else None,
                     "peak_source": peak_src,
                     "samples": int(k1_samples),
                     "note": "algorithmic bytes of the whole path (SURVEY 8-d: %.2f MB/img) / avg duration of the "
                             "dominant kernel (CUDA events around every %d-th launch of the timed region); see "
                             "roofline_step and kernels_ms" % (bytes_img / 1e6, K1_EVERY)},
        "roofline_step": {"achieved": path_gbs, "peak": peak, "unit": "GB/s", "frac": path_gbs / peak,
                          "bytes_per_image": bytes_img},
        "kernels_ms": kern,
        "host_issue_ms_per_step": host_issue_ms,
        "ms_per_step_gt_prepacked": ms_prepacked,
        "loss_stream_hbm_frac": (stream_bytes / (kern["loss_stream"] * 1e-3) / 1e9 / peak) if "loss_stream" in kern else None,
        "e2e": {"value": e2e_val, "unit": "images/s", "h2d_bytes_per_step": in_bytes + gt_rows_bytes,
                "d2h_bytes_per_step": 4, "steps": e_steps},
        # per step: k_pack_targets_mapped, k_gt_rects, k_gt_setup, k_cand_overlaps, k_topk_per_gt, k_resolve_image,
        # k_positive_gather, k_loss_stream_v4, k_loss_finalize, k_scale (torch's fills and copies are not counted)
        "fp16_inputs": {"note": "same workload with fp16 head outputs (autocast, the reference's default): maps read in place, "
                                "fp32 arithmetic, fp16 gradients",
                        "value": world * B / (ms_res16 / h_steps / 1e3), "ms_per_step": ms_res16 / h_steps,
                        "e2e": {"value": world * B / (ms_e2e16 / e_steps / 1e3), "unit": "images/s",
                                "h2d_bytes_per_step": in_bytes // 2 + gt_rows_bytes, "d2h_bytes_per_step": 4}},
        "dp_step": {"workload": f"configs[4]: Segment head of a yolov8x neck (cuDNN, fp16 autocast) + fused loss + backward, "
                                f"batch {B}/GPU, DistributedDataParallel over {world} rank(s), loss *= world_size",
                    "value": world * B / (ms_dp / 1e3), "unit": "images/s", "ms_per_step": ms_dp,
                    "ms_per_step_no_allreduce": ms_dp_nosync, "allreduce_exposed_ms": max(ms_dp - ms_dp_nosync, 0.0),
                    "allreduce_bytes": dp_params * 4, "collective": "NCCL all-reduce of the head gradients (DDP buckets)"},
        "gpu_launches": 10 * args.steps,
        "infer": {"metric": "decode+NMS images/sec", "value": inf_val, "unit": "images/s",
                  "workload": f"C3: batch {ib} @640, conf 0.25 / IoU 0.7, max_det 300, kept/img {kept:.0f}",
                  "ms_per_step": ms_inf / i_steps, "kernels_ms": ikern,
                  "roofline_frac": inf_bytes_img * ib / (ms_inf / i_steps * 1e-3) / 1e9 / peak,
                  "one_call_detect": {"note": "deployment form (ycr_detect): same rows, the prediction tensor is never written",
                                      "value": world * ib / (ms_det / i_steps / 1e3), "ms_per_step": ms_det / i_steps,
                                      "packed_output": {"note": "rows of all images back to back + device counts, as "
                                                                "ycr_detect returns them (no host read, no row views)",
                                                        "value": world * ib / (ms_det_packed / i_steps / 1e3),
                                                        "ms_per_step": ms_det_packed / i_steps}}},
        "clocks": clocks,
    }
    if cpu_val is not None:
        line["cpu_baseline"] = {"value": cpu_val, "unit": "images/s", "cores": cores, "kind": cpu_kind,
                                "sample": f"the first {args.cpu_images} images of this batch, fwd+bwd, 1 warm-up + 2 timed ("
                                          + ("the reference's v8SegmentationLoss, baseline/_ref" if cpu_kind == "reference"
                                             else "oracle/polar_oracle.seg_loss") + ")"}
    if per_rank is not None:
        line["per_rank"] = per_rank
    emit(line)
    if use_dist:
        import torch.distributed as dist
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2")
    ap.add_argument("--cpu-images", type=int, default=4)
    ap.add_argument("--quick", action="store_true", help="resident train-path timing only (for ncu runs)")
    ap.add_argument("--quick-infer", action="store_true", help="decode+NMS kernels only (for ncu runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
