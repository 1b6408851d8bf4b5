"""Multi-GPU tests (-m gpu, skipped on a one-GPU box): the two collectives of the port on NCCL over NVLink — the score gather of
evaluate() (evaluate.py:324, 396-417 in the reference) and the optional down-sampled retina hand-off — fed by
real kernel output of one simulator per rank.  Run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
    pytest.skip("needs two CUDA devices", allow_module_level=True)

import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from real_robots_b200.envs.vec_env import VecREALRobotEnv, gather_retinas
    from real_robots_b200.evaluate import gather_scores, shard_range
    n = 8
    env = VecREALRobotEnv(n, device=rank, objects=3, action_type='joints', additional_obs=True)
    env.reset()
    g = torch.Generator(device=dev); g.manual_seed(7)            # the same envs on every rank: results must not depend on the rank
    a = env.sample_random_actions(g)
    for _ in range(40):
        obs, _, _, _ = env.step({'joint_command': a, 'render': True})
    # score gather over NCCL: the goals are partitioned into contiguous shards (evaluate.shard_range)
    G = world * n - 3                                              # ragged last shard
    lo, hi = shard_range(G, rank, world)
    goal = obs['object_positions'].contiguous() + 0.05
    local = env.evaluate_goal(goal)[:hi - lo]
    vec = gather_scores(torch.arange(lo, hi, dtype=torch.int64, device=dev), local, G)
    # retina hand-off over NCCL: every rank / the learner only
    everywhere = gather_retinas(obs['retina'], factor=4)
    on_learner = gather_retinas(obs['retina'], factor=8, dst=0)
    q.put((rank, vec.cpu().numpy(), local.cpu().numpy(), everywhere.cpu().numpy(), None if on_learner is None else on_learner.cpu().numpy(),
           obs['retina'].cpu().numpy()))
    env.close()
    dist.destroy_process_group()


def test_score_gather_and_retina_handoff_over_nccl():
    world = 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(world):
        r = q.get(timeout=600)
        res[r[0]] = r[1:]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    vec0, loc0, every0, learner0, ret0 = res[0]
    vec1, loc1, every1, learner1, ret1 = res[1]
    assert (ret0 == ret1).all()                                    # an env's images do not depend on the GPU that hosts it
    G = 2 * 8 - 3
    assert vec0.shape == (G,) and (vec0 == vec1).all() and np.isfinite(vec0).all()
    from real_robots_b200.evaluate import shard_range
    (lo0, hi0), (lo1, hi1) = shard_range(G, 0, 2), shard_range(G, 1, 2)
    assert (lo0, hi1) == (0, G) and hi0 == lo1 and len(loc0) == hi0 - lo0 and len(loc1) == hi1 - lo1
    assert (vec0[lo0:hi0] == loc0).all() and (vec0[lo1:hi1] == loc1).all()
    small = ret0.reshape(8, 60, 4, 80, 4, 3).astype(np.float32).mean(axis=(2, 4)).round().clip(0, 255).astype(np.uint8)
    assert every0.shape == (16, 60, 80, 3) and (every0 == every1).all() and (every0[:8] == small).all() and (every0[8:] == small).all()
    assert learner1 is None and learner0.shape == (16, 30, 40, 3)
