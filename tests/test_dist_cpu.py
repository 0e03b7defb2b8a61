"""CPU, world_size 2, gloo: the N>1 host logic (sharding + the disparity all-gather)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from disprcnn_b200.parallel import gather_disparity, shard_range, sharded_forward, sharded_forward_async


def test_shard_range_partitions():
    for B in (0, 1, 5, 32, 33, 255, 256):
        for G in (1, 2, 3, 8):
            spans = [shard_range(B, r, G) for r in range(G)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(G - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


class _FakeModel:
    """Stands in for PSMNet.forward_features: a per-ROI function, so gathering is checkable."""

    def forward_features(self, l, r, H=None, W=None):
        return (l.sum(1) - r.sum(1)) * 0.5


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(3)
        L = torch.randn(B, 4, 6, 7, generator=g)
        R = torch.randn(B, 4, 6, 7, generator=g)
        full = sharded_forward(_FakeModel(), L, R)
        want = _FakeModel().forward_features(L, R)
        lo, hi = shard_range(B, rank, world)
        local = sharded_forward(_FakeModel(), L, R, gather=False)
        ok = torch.equal(full, want) and torch.equal(local, want[lo:hi]) and full.shape[0] == B
        if B % world == 0:  # the overlapped form: sub-chunked async gathers, two batches in flight before the first wait()
            for chunks in (1, 2, 3):
                ok = ok and torch.equal(sharded_forward(_FakeModel(), L, R, chunks=chunks), want)
            p1 = sharded_forward_async(_FakeModel(), L, R, chunks=2)
            p2 = sharded_forward_async(_FakeModel(), L * 2, R * 2, chunks=2)
            ok = ok and torch.equal(p1.wait(), want) and torch.equal(p2.wait(), want * 2) and p1.wait() is p1.wait()
            p3 = sharded_forward_async(_FakeModel(), L[lo:hi], R[lo:hi], chunks=2, presharded=True)
            ok = ok and torch.equal(p3.wait(), want)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('B', [8, 5])  # equal shards and ragged shards
def test_gather_world2_gloo(B):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
    assert res == [(0, True), (1, True)]


def test_gather_without_process_group_is_identity():
    x = torch.randn(3, 4, 5)
    assert gather_disparity(x, 3) is x


def _pred_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from disprcnn_b200.parallel import gather_predictions

        def make(img):
            g = torch.Generator().manual_seed(100 + img)
            n = img % 3   # some images have no detections
            return {'bbox': torch.rand(n, 4, generator=g) * 300, 'scores': torch.rand(n, generator=g), 'labels': torch.randint(0, 4, (n,), generator=g),
                    'disparity': torch.rand(5 + img, 7, generator=g), 'keep': torch.rand(n, generator=g) > 0.5,
                    'mask': (torch.rand(n, 3, 3, generator=g) * 255).to(torch.uint8)}
        mine = {img: make(img) for img in range(7) if img % world == rank}     # images dealt round-robin, like the eval sampler
        every = gather_predictions(mine)
        only0 = gather_predictions(mine, dst=0)
        ok = list(every) == list(range(7)) and (only0 is None) == (rank != 0)
        for img in range(7):
            want = make(img)
            for k, v in want.items():
                got = every[img][k]
                ok = ok and got.dtype == v.dtype and tuple(got.shape) == tuple(v.shape) and torch.equal(got, v)
        empty = gather_predictions({} if rank == 1 else {3: make(3)})            # a rank with nothing to contribute
        ok = ok and list(empty) == [3]
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_typed_prediction_gather_world2_gloo():
    """per-image prediction fields travel as typed raw bytes in one padded all-gather (no pickling)."""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_pred_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
    assert res == [(0, True), (1, True)]


def test_typed_prediction_gather_without_process_group():
    from disprcnn_b200.parallel import gather_predictions
    p = {4: {'a': torch.ones(2)}, 1: {'a': torch.zeros(3)}}
    out = gather_predictions(p)
    assert list(out) == [1, 4] and out[4]['a'] is p[4]['a']
