"""Data-parallel plumbing of the path: every image is an independent unit, so the
batch is sharded by contiguous images per rank (what DistributedSampler + `batch // world_size` do in
the reference, engine/trainer.py:241) and the hot path itself needs no data-path collective.  Reference
DDP semantics are kept: each rank normalises by its LOCAL target_scores_sum and `loss *= world_size`
(engine/trainer.py:365).  `global_target_scores_sum` is the optional one-scalar all-reduce."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(batch_size: int, rank: int, world: int):
    """Contiguous image range [lo, hi) of `rank`; sizes differ by at most one."""
    base, rem = divmod(batch_size, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(batch: dict, batch_size: int, rank: int, world: int) -> dict:
    """Rows of the collated `batch` dict that belong to this rank's images, image index re-based to 0."""
    lo, hi = shard_range(batch_size, rank, world)
    bi = batch["batch_idx"].view(-1)
    keep = (bi >= lo) & (bi < hi)
    segs = batch["segments"]
    out = {
        "batch_idx": bi[keep] - lo,
        "cls": batch["cls"][keep],
        "bboxes": batch["bboxes"][keep],
        "segments": list(segs[lo:hi]) if isinstance(segs, (list, tuple)) else segs[keep],
    }
    if "img" in batch:
        out["img"] = batch["img"][lo:hi]
    return out


def max_over_ranks(values, device=None):
    """Element-wise max of a few host floats over all ranks (device timings: max over ranks)."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t.tolist()]


def global_target_scores_sum(local_sum: torch.Tensor) -> torch.Tensor:
    """Optional global-normaliser mode: one fp32 scalar all-reduce (SUM)."""
    t = local_sum.clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def scale_loss_for_ddp(loss: torch.Tensor, world: int) -> torch.Tensor:
    """engine/trainer.py:365 — DDP averages gradients, the reference multiplies the loss back."""
    return loss * world if world > 1 else loss


def rescale_to_global_norm(total: torch.Tensor, items: torch.Tensor, local_tss: torch.Tensor):
    """Optional global-normaliser mode of v8SegmentationLoss: the reference normalises each rank's
    loss by its LOCAL max(target_scores.sum(), 1) (utils/loss.py:866) and lets DDP average the gradients; with
    `global_norm` every rank divides by the sum over ranks instead - one fp32 scalar all-reduce, and the loss (hence,
    through autograd, every gradient) is rescaled by local / global.  `local_tss` is the clamped local normaliser
    the kernels used (exact whenever every rank has a target-score sum >= 1)."""
    glob = global_target_scores_sum(local_tss.detach().float()).clamp(min=1.0)
    ratio = (local_tss.detach().float() / glob).to(total.dtype)
    return total * ratio, items * ratio
