"""The multi-GPU sharder's host logic on CPU: world_size-2 (and 3) `gloo` process groups, uneven batches, ranks
without work, B_ref propagation and input-order reassembly.  The restore function is a stand-in (the kernels need
a GPU); on the GPU box the same code path runs with backend nccl (bench.py --gpus N, Defender.defend_point_cloud_sharded)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ifdefense_b200 import shard


def fake_restore(base):
    def fn(lo, hi, B_ref):                       # depends on the cloud, its global index and the batch's B_ref
        return base[lo:hi] * 2.0 + np.arange(lo, hi, dtype=np.float32)[:, None, None] + 1000.0 * B_ref
    return fn


def expected(base, n, bs):
    out = np.zeros_like(base)
    for lo, hi in shard.batch_bounds(n, bs):
        out[lo:hi] = fake_restore(base)(lo, hi, hi - lo)
    return out


def _worker(rank, world, port, n, bs, K, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        base = np.random.default_rng(0).normal(size=(n, K, 3)).astype(np.float32)
        got = shard.restore_sharded(fake_restore(base), n, bs)
        q.put((rank, np.array_equal(got, expected(base, n, bs)), got.shape))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_plan_covers_every_cloud_once_with_reference_batches():
    for n, bs, w in ((2468, 192, 8), (10, 4, 3), (5, 192, 8), (1, 1, 2), (64, 64, 1), (13, 5, 4)):
        segs = shard.plan(n, bs, w)
        seen = np.zeros(n, int)
        for r in range(w):
            for a, b, B_ref in segs[r]:
                seen[a:b] += 1
                lo = (a // bs) * bs
                assert B_ref == min(lo + bs, n) - lo and lo <= a < b <= lo + B_ref
        assert (seen == 1).all()
        per = [sum(b - a for a, b, _ in s) for s in segs]
        assert max(per) - min(per) <= len(shard.batch_bounds(n, bs))          # one cloud per batch at most
    assert shard.batch_bounds(2468, 192)[-1] == (2304, 2468)                 # the reference's ragged last batch (164)


def test_single_process_is_the_identity_plan():
    base = np.random.default_rng(1).normal(size=(7, 4, 3)).astype(np.float32)
    got = shard.restore_sharded(fake_restore(base), 7, 3, rank=0, world=1)
    assert np.array_equal(got, expected(base, 7, 3))


@pytest.mark.parametrize("world,n,bs", [(2, 11, 4), (2, 1, 4), (3, 8, 3)])
def test_gloo_world(world, n, bs):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, bs, 5, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r for r, _, _ in res) == list(range(world))
    assert all(ok for _, ok, _ in res) and all(shape == (n, 5, 3) for _, _, shape in res)


def _mesh_stub(generator, encode_inputs, pc, num_points=1024, rng=None):
    """Stand-in for mesh.resample_points (needs a GPU): depends on the cloud and on the cloud's own random stream."""
    return (pc[:num_points] * 0.5 + rng.random((num_points, 3))).astype(np.float32)


def _mesh_worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ifdefense_b200 import mesh
        clouds = np.random.default_rng(3).normal(size=(n, 16, 3)).astype(np.float32)
        got = mesh.resample_points_sharded(None, None, clouds, num_points=8, seed=5, resample=_mesh_stub)
        want = mesh.resample_points_sharded(None, None, clouds, num_points=8, seed=5, resample=_mesh_stub, rank=0, world=1)
        q.put((rank, np.array_equal(got, want), got.shape))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 5), (3, 2)])
def test_gloo_mesh_resampling_is_independent_of_the_world_size(world, n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_mesh_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok and shape == (n, 8, 3) for _, ok, shape in res)
