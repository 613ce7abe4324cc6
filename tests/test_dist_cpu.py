"""N>1 path on CPU (gloo, world_size 2): the one-time parameter broadcast delivers a bit-identical
pack to every rank, batch sharding covers the batch exactly once, logits gather restores order."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _make_pack():
    from ivit_b200.pack import Pack
    rng = np.random.default_rng(0)
    arrays = {"a.weight_integer": rng.integers(-128, 128, (37, 48)).astype(np.int8),
              "a.bias_integer": rng.integers(-2 ** 20, 2 ** 20, 37).astype(np.int32),
              "q.me": rng.integers(2 ** 30, 2 ** 31 - 1, (5, 2)).astype(np.int32),
              "pos": rng.integers(-30000, 30000, (7, 3)).astype(np.int16),
              "s": rng.random(3).astype(np.float32)}
    return Pack({"arch": "deit", "embed_dim": 48}, arrays)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ivit_b200.dist import broadcast_pack, gather_logits, shard_batch
        ref = _make_pack()
        got = broadcast_pack(ref if rank == 0 else None, src=0, device="cpu")
        ok = got.meta == ref.meta and set(got.arrays) == set(ref.arrays)
        for k in ref.arrays:
            ok &= bool(np.array_equal(got[k], ref[k])) and got[k].dtype == ref[k].dtype
        lo, hi = shard_batch(11, rank, world)
        local = torch.arange(lo, hi, dtype=torch.float32).reshape(-1, 1).repeat(1, 4)
        full = gather_logits(local, 11)
        ok &= bool(torch.equal(full[:, 0], torch.arange(11, dtype=torch.float32)))
        q.put((rank, ok, (lo, hi)))
    finally:
        dist.destroy_process_group()


def test_broadcast_and_sharding_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    spans = sorted(s for _, _, s in res)
    assert spans == [(0, 6), (6, 11)]


def test_blob_roundtrip_alignment():
    from ivit_b200.dist import blob_to_pack, pack_to_blob
    p = _make_pack()
    man, blob = pack_to_blob(p)
    assert all(off % 256 == 0 for (_, _, _, off, _) in man["entries"])
    p2 = blob_to_pack(man, blob)
    for k in p.arrays:
        assert np.array_equal(p2[k], p[k])


@pytest.mark.parametrize("n,world", [(256, 8), (10, 4), (3, 8), (2048, 8)])
def test_shard_batch_partitions(n, world):
    from ivit_b200.dist import shard_batch
    spans = [shard_batch(n, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
