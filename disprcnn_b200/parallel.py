"""ROI-batch sharding across the GPUs of one node + the single all-gather of disparity maps.

ROIs are independent in eval (no cross-ROI op in stackhourglass.py:106-174; BatchNorm uses running
statistics), so the path shards with NO data-path collective: rank r owns the contiguous chunk
``shard_range(B, r, G)`` of the ROI batch, weights (7.5 MB) are replicated.  The one exchange is
the gather of the per-ROI disparity maps [B/G,H,W] f32 -> [B,H,W] on every rank -- the typed
replacement of the reference's pickled ``all_gather`` of BoxLists carrying 'disparity' fields
(disprcnn/utils/comm.py:47-87, disprcnn/engine/inference.py:53-72).  One process per GPU,
``torch.distributed`` (NCCL over NVLink/NVSwitch on the GPUs, gloo for the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_range(B, rank, world):
    """Contiguous chunk [lo, hi) of B ROIs owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(B, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_disparity(local, B, group=None):
    """All-gather the per-rank disparity maps into the full [B,H,W] tensor (same on every rank).

    `local` is this rank's [hi-lo, H, W] block.  Equal shards use one ``all_gather_into_tensor``
    (a single NCCL all-gather writing straight into the output); ragged shards pad to the
    largest shard first.
    """
    if not dist.is_available() or not dist.is_initialized():
        assert local.shape[0] == B
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    H, W = local.shape[1:]
    lo, hi = shard_range(B, rank, world)
    assert local.shape[0] == hi - lo, f'rank {rank}: expected {hi - lo} ROIs, got {local.shape[0]}'
    local = local.contiguous()
    if B % world == 0:
        out = torch.empty((B, H, W), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local, group=group)
        return out
    mx = -(-B // world)
    pad = torch.zeros((mx, H, W), dtype=local.dtype, device=local.device)
    pad[:hi - lo] = local
    buf = torch.empty((world * mx, H, W), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf, pad, group=group)
    parts = []
    for r in range(world):
        a, b = shard_range(B, r, world)
        parts.append(buf[r * mx:r * mx + (b - a)])
    return torch.cat(parts, 0)


def sharded_forward(model, left_fea, right_fea, H=None, W=None, group=None, gather=True):
    """Run `model.forward_features` on this rank's shard of the (replicated or pre-sharded) batch.

    left_fea/right_fea: the FULL [B,C,Hf,Wf] batch (each rank slices its own chunk).  Returns the
    gathered [B,H,W] disparity (or the local block when gather=False).
    """
    B = left_fea.shape[0]
    if dist.is_available() and dist.is_initialized():
        lo, hi = shard_range(B, dist.get_rank(group), dist.get_world_size(group))
    else:
        lo, hi = 0, B
    local = model.forward_features(left_fea[lo:hi], right_fea[lo:hi], H, W)
    return gather_disparity(local, B, group) if gather else local
