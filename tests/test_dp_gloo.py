"""CPU, world size 2 over gloo: the host-side data-parallel logic (window sharding, flat-gradient all-reduce
with the unused-parameter tail excluded, metric partial-sum reduction)."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import pcm_b200  # noqa: F401
from pcm_b200 import parallel


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # flat gradient: first 10 entries are reduced, the tail (never-used params) is left alone
        g = torch.arange(14, dtype=torch.float32) * (rank + 1)
        scale = parallel.allreduce_flat_grads(g, 10)
        # metric partial sums over time shards
        part = torch.full((2, 3, 4, 5), float(rank + 1), dtype=torch.float64)
        part, t_total = parallel.allreduce_metric_partials(part, 7 + rank)
        idx = parallel.shard_windows(8109, rank, world, seed=42, epoch=3)
        out.put((rank, g.numpy(), scale, part.numpy().copy(), t_total, idx.numpy()))
    finally:
        dist.destroy_process_group()


def test_dp_world2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    base = np.arange(14, dtype=np.float32)
    for rank, g, scale, part, t_total, idx in res:
        np.testing.assert_allclose(g[:10], base[:10] * 3)              # summed over ranks
        np.testing.assert_allclose(g[10:], base[10:] * (rank + 1))     # tail untouched
        assert scale == 0.5
        np.testing.assert_allclose(part, 3.0)
        assert t_total == 15
    i0, i1 = res[0][5], res[1][5]
    assert len(i0) == len(i1) == (8109 + 1) // 2
    allidx = np.concatenate([i0, i1])
    assert set(allidx.tolist()) == set(range(8109))                     # every window visited
    assert len(allidx) - len(set(allidx.tolist())) == 1                 # one wrap-around pad, like DistributedSampler


def test_shard_windows_matches_torch_distributed_sampler():
    from torch.utils.data.distributed import DistributedSampler
    ds = list(range(1001))
    for world in (2, 4, 8):
        for rank in range(world):
            s = DistributedSampler(ds, num_replicas=world, rank=rank, shuffle=True, seed=42)
            s.set_epoch(5)
            ours = parallel.shard_windows(1001, rank, world, seed=42, epoch=5).tolist()
            assert ours == list(iter(s))


def test_single_process_is_identity():
    g = torch.arange(6, dtype=torch.float32)
    assert parallel.allreduce_flat_grads(g, 4) == 1.0
    np.testing.assert_allclose(g.numpy(), np.arange(6))
