"""N > 1 host logic on CPU (gloo, world_size 2): the point sharding is an exact partition,
the sum over ranks of the per-shard camera-side gradients equals the full gradient (what
the per-step all-reduce computes), and merge_points reassembles the model."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from particlesfm_b200 import distributed, synthetic as syn
    prob, _ = syn.make_ba_problem(9, 700, 5, seed=13, track_len_range=(2, 8))
    o = oracle.ba_global_options(True, True)
    shard = prob.shard(rank, world)
    # exact partition of the observations, whole points per rank
    cnt = torch.tensor([shard.num_observations], dtype=torch.int64)
    dist.all_reduce(cnt)
    own = torch.zeros(prob.num_points, dtype=torch.int64)
    own[distributed.owned_points(shard)] = 1
    dist.all_reduce(own)
    c, r, gc, gp = oracle.ba_evaluate(shard, o, num_threads=2)
    tc, tg, tp = torch.tensor([c], dtype=torch.float64), torch.from_numpy(gc.copy()), torch.from_numpy(gp.copy())
    dist.all_reduce(tc); dist.all_reduce(tg); dist.all_reduce(tp)
    c0, _, gc0, gp0 = oracle.ba_evaluate(prob, o, num_threads=2)
    # merge_points: every rank perturbs only its own points, the merge yields all of them
    mine = distributed.owned_points(shard)
    local = shard.copy()
    local.xyz[mine] += 1.0 + rank
    distributed.merge_points(local, dist, world)
    expect = prob.xyz.copy()
    for rk in range(world):
        expect[distributed.owned_points(prob.shard(rk, world))] += 1.0 + rk
    res = dict(part=int(cnt.item()) == prob.num_observations, own=bool((own[np.bincount(prob.obs_point, minlength=prob.num_points) > 0] == 1).all()),
               cost=abs(tc.item() - c0) <= 1e-10 * c0, gc=float(np.abs(tg.numpy() - gc0).max() / np.abs(gc0).max()),
               gp=float(np.abs(tp.numpy() - gp0).max() / np.abs(gp0).max()), merge=bool(np.array_equal(local.xyz, expect)))
    q.put((rank, res))
    dist.barrier()
    dist.destroy_process_group()


def test_sharding_and_allreduce_semantics_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29400 + os.getpid() % 500
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, res in out:
        assert res["part"] and res["own"] and res["cost"] and res["merge"], (rank, res)
        assert res["gc"] < 1e-12 and res["gp"] < 1e-12, (rank, res)


def test_shard_balance():
    sys.path.insert(0, ROOT)
    from particlesfm_b200 import synthetic as syn
    prob, _ = syn.make_ba_problem(20, 5000, 8, seed=3, track_len_range=(2, 20))
    for world in (2, 4, 8):
        sizes = [prob.shard(r, world).num_observations for r in range(world)]
        assert sum(sizes) == prob.num_observations
        assert max(sizes) - min(sizes) <= 2 * 20          # balanced by observation count, whole points
