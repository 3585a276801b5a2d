"""CPU, world_size 2 over gloo: the host-side data-parallel logic of the path (sharding by image,
max-over-ranks timing reduction, the optional one-scalar normaliser all-reduce)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from util import ROOT  # noqa: F401
import ycr_b200  # noqa: F401
from ycr_b200 import synth, dp
from oracle import polar_oracle as po


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cfg = synth.PathConfig("dp", 5, 4, 160, nc=10)            # odd batch: ranks get 3 and 2 images
        batch = synth.make_gts(cfg, 77, ragged=True)
        feats = synth.make_feats(cfg, 77)
        lo, hi = dp.shard_range(cfg.batch, rank, world)
        sub = dp.shard_batch(batch, cfg.batch, rank, world)
        assert int(sub["batch_idx"].max().item() if sub["batch_idx"].numel() else 0) < hi - lo
        assert len(sub["segments"]) == hi - lo
        # the shard's packed targets equal the corresponding rows of the full batch's packing
        full = po.pack_targets(batch, cfg.batch, (160, 160))
        mine = po.pack_targets(sub, hi - lo, (160, 160))
        g = mine.shape[1]
        assert torch.equal(mine, full[lo:hi, :g]) and not bool(full[lo:hi, g:].any())
        # rank-local loss with LOCAL normaliser (reference DDP semantics), then the global option
        r = po.seg_loss([f[lo:hi] for f in feats], sub, cfg.strides, cfg.nc, cfg.rays, with_grad=False)
        local = torch.tensor(float(r["assign"]["target_scores"].sum()))
        glob = dp.global_target_scores_sum(local)
        ref = po.seg_loss(feats, batch, cfg.strides, cfg.nc, cfg.rays, with_grad=False)
        assert abs(float(glob) - float(ref["assign"]["target_scores"].sum())) < 1e-4 * max(1.0, float(glob))
        # image independence: per-rank assignment == slice of the full-batch assignment
        assert torch.equal(r["assign"]["target_gt_idx"], ref["assign"]["target_gt_idx"][lo:hi])
        assert torch.equal(r["assign"]["fg_mask"], ref["assign"]["fg_mask"][lo:hi])
        # global-normaliser mode: every rank's rescaled loss numerators add up to the full-batch loss
        tss_local = torch.tensor(max(float(local), 1.0))
        tot, items = dp.rescale_to_global_norm(r["loss"] / (hi - lo), r["loss_items"], tss_local)
        both = items.clone()
        dist.all_reduce(both)
        if float(ref["assign"]["target_scores"].sum()) >= 1.0 and float(local) >= 1.0:
            assert float((both - ref["loss_items"]).abs().max()) <= 2e-5 * float(ref["loss_items"].abs().max())
        ms = dp.max_over_ranks([10.0 + rank, 5.0 - rank])
        assert ms == [10.0 + world - 1, 5.0]
        assert float(dp.scale_loss_for_ddp(torch.tensor(2.0), world)) == 2.0 * world
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_shard_ranges_cover_batch():
    for B in (1, 5, 64, 512):
        for w in (1, 2, 4, 8):
            spans = [dp.shard_range(B, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
