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


class PendingGather:
    """An all-gather in flight on the collective's own (side) stream.  ``wait()`` makes the CURRENT stream wait for it and
    returns the gathered [B,H,W] tensor; until then the current stream is free to run the next batch's kernels."""

    def __init__(self, works, bufs, B, world, chunk_sizes, shard):
        self._works, self._bufs, self._B, self._world, self._chunks, self._shard = works, bufs, B, world, chunk_sizes, shard
        self._out = None

    def wait(self):
        if self._out is not None:
            return self._out
        for w in self._works:
            w.wait()   # stream-level wait (no host block): current stream <- the NCCL stream's completion event
        if len(self._bufs) == 1:
            self._out = self._bufs[0].view(self._B, *self._bufs[0].shape[2:])
        else:
            # chunk c of rank r sits at bufs[c][r]; the full batch is rank-major: [r][c0 | c1 | ...]
            self._out = torch.cat([torch.cat([b[r] for b in self._bufs], 0) for r in range(self._world)], 0)
        return self._out


def sharded_forward_async(model, left_fea, right_fea, H=None, W=None, group=None, chunks=1, presharded=False):
    """Shard, compute, and START the gather without blocking the compute stream ("issue on a side stream in
    >= 2 sub-chunks to overlap with the tail ROIs' compute").

    The local shard is processed in ``chunks`` ROI sub-chunks; as soon as a sub-chunk's soft-argmin is enqueued its
    ``all_gather_into_tensor`` is issued with ``async_op=True`` -- NCCL runs it on its own stream behind an event, so the
    gather of sub-chunk c overlaps the kernels of sub-chunk c+1 (and, when the caller defers ``wait()``, of the next batch).
    Requires equal shards (B a multiple of the world size; pad the batch otherwise -- ``sharded_forward`` handles ragged
    batches with the blocking path).  Returns a ``PendingGather``.
    """
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if presharded:
        B = left_fea.shape[0] * world
        lo, hi = 0, left_fea.shape[0]
    else:
        B = left_fea.shape[0]
        lo, hi = shard_range(B, rank, world)
    if B % world:
        raise RuntimeError(f'sharded_forward_async: batch {B} is not a multiple of the world size {world}')
    per = B // world
    chunks = max(1, min(int(chunks), per))
    sizes = [per // chunks + (1 if c < per % chunks else 0) for c in range(chunks)]
    works, bufs, a = [], [], lo
    for n in sizes:
        local = model.forward_features(left_fea[a:a + n], right_fea[a:a + n], H, W).contiguous()
        buf = torch.empty((world, n) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        works.append(dist.all_gather_into_tensor(buf.view(world * n, *local.shape[1:]), local, group=group, async_op=True))
        bufs.append(buf)
        a += n
    return PendingGather(works, bufs, B, world, sizes, (lo, hi))


def sharded_forward(model, left_fea, right_fea, H=None, W=None, group=None, gather=True, chunks=1):
    """Run `model.forward_features` on this rank's shard of the (replicated or pre-sharded) batch.

    left_fea/right_fea: the FULL [B,C,Hf,Wf] batch (each rank slices its own chunk).  Returns the
    gathered [B,H,W] disparity (or the local block when gather=False).  ``chunks`` > 1 (equal shards only) overlaps the
    gather of each ROI sub-chunk with the next sub-chunk's kernels (see ``sharded_forward_async``).
    """
    B = left_fea.shape[0]
    if dist.is_available() and dist.is_initialized():
        world = dist.get_world_size(group)
        lo, hi = shard_range(B, dist.get_rank(group), world)
        if gather and chunks > 1 and B % world == 0:
            return sharded_forward_async(model, left_fea, right_fea, H, W, group, chunks).wait()
    else:
        lo, hi = 0, B
    local = model.forward_features(left_fea[lo:hi], right_fea[lo:hi], H, W)
    return gather_disparity(local, B, group) if gather else local


# ---------------------------------------------------------------------------------------------------------------------
# Typed gather of per-image predictions
# ---------------------------------------------------------------------------------------------------------------------
_DTYPES = [torch.float32, torch.float64, torch.float16, torch.bfloat16, torch.int64, torch.int32, torch.int16, torch.int8, torch.uint8,
           torch.bool]


def _pack_predictions(preds, device):
    """{image_id: {field: tensor}} -> (int64 header, uint8 payload).  Header layout (all int64):
    [n_images, n_fields, (len(name), name bytes...) per field, then per image: image_id, per field: dtype code, ndim, dims...]."""
    ids = sorted(preds)
    names = sorted(preds[ids[0]]) if ids else []
    head = [len(ids), len(names)]
    for n in names:
        b = n.encode()
        head += [len(b)] + list(b)
    chunks = []
    for i in ids:
        if sorted(preds[i]) != names:
            raise RuntimeError(f'gather_predictions: image {i} has fields {sorted(preds[i])}, expected {names}')
        head.append(int(i))
        for n in names:
            t = preds[i][n]
            if t.dtype not in _DTYPES:
                raise RuntimeError(f'gather_predictions: dtype {t.dtype} of field {n!r} is not supported')
            head += [_DTYPES.index(t.dtype), t.dim()] + list(t.shape)
            flat = t.detach().to(device).contiguous().reshape(-1)
            chunks.append((flat.view(torch.uint8) if flat.dtype != torch.bool else flat.to(torch.uint8)).reshape(-1))
            pad = (-chunks[-1].numel()) % 8   # keep every field 8-byte aligned inside the payload
            if pad:
                chunks.append(torch.zeros(pad, dtype=torch.uint8, device=device))
    payload = torch.cat(chunks) if chunks else torch.zeros(0, dtype=torch.uint8, device=device)
    return torch.tensor(head, dtype=torch.int64, device=device), payload


def _unpack_predictions(head, payload, out):
    head = head.tolist()
    n_img, n_fields, p = head[0], head[1], 2
    names = []
    for _ in range(n_fields):
        ln = head[p]
        names.append(bytes(head[p + 1:p + 1 + ln]).decode())
        p += 1 + ln
    off = 0
    for _ in range(n_img):
        img = head[p]
        p += 1
        fields = {}
        for n in names:
            dt, nd = _DTYPES[head[p]], head[p + 1]
            shape = head[p + 2:p + 2 + nd]
            p += 2 + nd
            numel = 1
            for s in shape:
                numel *= s
            nbytes = numel * (1 if dt == torch.bool else torch.empty(0, dtype=dt).element_size())
            raw = payload[off:off + nbytes]
            fields[n] = (raw.to(torch.bool) if dt == torch.bool else raw.view(dt)).reshape(shape)
            off += nbytes + ((-nbytes) % 8)
        out[img] = fields
    return out


def gather_predictions(predictions, group=None, dst=None):
    """Typed replacement of the reference's pickled prediction gather (disprcnn/engine/inference.py:53-72
    ``_accumulate_predictions_from_multiple_gpus`` -> disprcnn/utils/comm.py:47-87 ``all_gather``: every rank pickles its
    ``{image_id: BoxList}`` dict -- BoxLists that carry image-sized float 'disparity' maps -- into a ByteTensor through the
    host).  Here ``predictions`` is ``{image_id: {field name: tensor}}`` (e.g. 'bbox' [R,4] f32, 'scores' [R] f32, 'labels' [R]
    i64, 'disparity' [H,W] f32); field tensors travel as raw bytes of their own dtype in ONE padded ``all_gather_into_tensor``
    (plus one small int64 header gather) -- NCCL over NVLink on the GPUs, no pickling, no host staging of the payload.
    Returns the merged dict ordered by image id -- like the reference, only where it is needed: on rank ``dst`` (None elsewhere),
    or on every rank when ``dst`` is None.  Without an initialised process group it returns ``predictions`` sorted by id."""
    if not dist.is_available() or not dist.is_initialized():
        return {i: predictions[i] for i in sorted(predictions)}
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    any_t = next((t for f in predictions.values() for t in f.values()), None)
    backend = dist.get_backend(group)
    device = torch.device('cuda', torch.cuda.current_device()) if backend == 'nccl' else torch.device('cpu')
    if any_t is not None and backend != 'nccl':
        device = torch.device('cpu')
    head, payload = _pack_predictions(predictions, device)
    sizes = torch.tensor([head.numel(), payload.numel()], dtype=torch.int64, device=device)
    all_sizes = torch.empty((world, 2), dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(all_sizes.view(-1), sizes, group=group)
    all_sizes = all_sizes.cpu()
    mh, mp = int(all_sizes[:, 0].max()), (max(int(all_sizes[:, 1].max()), 8) + 7) // 8 * 8
    hbuf = torch.zeros(mh, dtype=torch.int64, device=device)
    hbuf[:head.numel()] = head
    pbuf = torch.zeros(mp, dtype=torch.uint8, device=device)
    pbuf[:payload.numel()] = payload
    heads = torch.empty(world * mh, dtype=torch.int64, device=device)
    pays = torch.empty(world * mp, dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(heads, hbuf, group=group)
    dist.all_gather_into_tensor(pays, pbuf, group=group)
    if dst is not None and rank != dst:
        return None
    heads = heads.cpu()
    merged = {}
    for r in range(world):
        nh, npay = int(all_sizes[r, 0]), int(all_sizes[r, 1])
        _unpack_predictions(heads[r * mh:r * mh + nh], pays[r * mp:r * mp + npay], merged)
    return {i: merged[i] for i in sorted(merged)}
