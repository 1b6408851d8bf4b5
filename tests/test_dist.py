"""world_size-2 gloo test (CPU) of the only collective near the path: the all_gather of per-goal scores
(evaluate.py gather_scores), with goals partitioned across ranks (no data-path collective)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from real_robots_b200.evaluate import gather_scores, shard_range


def _worker(rank, world, port, n_items, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n_items, rank, world)
    ids = torch.arange(lo, hi, dtype=torch.int64)
    scores = ids.to(torch.float64) * 0.5 + 1.0          # score of goal g is a pure function of g
    out = gather_scores(ids, scores, n_items)
    q.put((rank, out.numpy()))
    dist.destroy_process_group()


def test_gather_scores_gloo_world2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n_items = 7                                         # ragged: 4 + 3
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_items, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = np.arange(n_items) * 0.5 + 1.0
    assert np.allclose(res[0], expect) and np.allclose(res[1], expect)


def test_gather_scores_without_process_group():
    ids = torch.tensor([2, 0])
    out = gather_scores(ids, torch.tensor([5.0, 7.0], dtype=torch.float64), 3)
    assert out[0] == 7 and out[2] == 5 and torch.isnan(out[1])


def _retina_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from real_robots_b200.envs.vec_env import gather_retinas
    g = torch.Generator().manual_seed(100 + rank)
    retina = torch.randint(0, 256, (3, 240, 320, 3), dtype=torch.uint8, generator=g)
    everywhere = gather_retinas(retina, factor=4)
    on_learner = gather_retinas(retina, factor=8, dst=0)
    q.put((rank, everywhere.numpy(), None if on_learner is None else on_learner.numpy()))
    dist.destroy_process_group()


def test_gather_retinas_gloo_world2():
    """The optional observation hand-off after the path: down-sampled retinas of both ranks on every rank / on the learner."""
    from real_robots_b200.envs.vec_env import gather_retinas
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_retina_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {r: (a, b) for r, a, b in (q.get(timeout=120) for _ in range(2))}
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    local = [gather_retinas(torch.randint(0, 256, (3, 240, 320, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(100 + r)), 4)
             for r in range(2)]
    expect = torch.cat(local, dim=0).numpy()
    assert res[0][0].shape == (6, 60, 80, 3) and (res[0][0] == expect).all() and (res[1][0] == expect).all()
    assert res[0][1].shape == (6, 30, 40, 3) and res[1][1] is None
