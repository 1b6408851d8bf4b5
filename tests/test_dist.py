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


class _FakeVecEnv:
    """CPU stand-in for VecREALRobotEnv inside evaluate_vectorised: objects move by the action every step, the score is
    the reference's formula (env.py:181-200) — enough to check that goal g is run with ITS poses on the rank that owns
    it and that its score lands in slot g on every rank."""
    made = []

    def __init__(self, n_envs, device=0, objects=3, action_type='joints', additional_obs=True, **cfg):
        self.n_envs, self.device, self.n_objects = n_envs, torch.device("cpu"), objects
        self.extrinsic_timesteps, self.goal_idx, self.goal_positions = 0, -1, None
        self.pos = torch.zeros(n_envs, objects, 3, dtype=torch.float64)
        self.steps = 0
        self._all = torch.ones(n_envs, dtype=torch.uint8)
        _FakeVecEnv.made.append(self)

    def reset(self, mask=None, render=True):
        return {"object_positions": self.pos}

    def _observe(self, render_mask):
        return {"object_positions": self.pos.clone(), "rendered": render_mask is not None}

    def set_object_poses(self, pos_quat, mask=None):
        self.pos = pos_quat[:, :, :3].clone()

    def step(self, action):
        self.pos = self.pos + action["shift"][:, None, :]
        self.steps += 1
        done = torch.full((self.n_envs,), self.steps >= self.extrinsic_timesteps)
        return {"object_positions": self.pos}, torch.zeros(self.n_envs, dtype=torch.float64), done, {}

    def evaluate_goal(self, goal_xyz):
        d = (self.pos - goal_xyz).norm(dim=2)
        return torch.exp(-(-np.log(0.25) / 0.10) * d).sum(dim=1)

    def close(self):
        pass


def _goals(n_goals, n_obj=2):
    rng = np.random.default_rng(5)
    init = np.zeros((n_goals, n_obj, 7)); init[:, :, 6] = 1
    init[:, :, :3] = rng.uniform(-0.3, 0.3, (n_goals, n_obj, 3))
    final = init[:, :, :3] + rng.uniform(-0.1, 0.1, (n_goals, n_obj, 3))
    return {"initial": init, "final": final}


def _expected_scores(goals, steps):
    pos = goals["initial"][:, :, :3] + steps * np.array([0.001, 0.0, 0.0])
    return np.exp(-(-np.log(0.25) / 0.10) * np.linalg.norm(pos - goals["final"], axis=2)).sum(1)


def _eval_worker(rank, world, port, n_goals, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    import real_robots_b200.envs.vec_env as ve
    from real_robots_b200.evaluate import evaluate_vectorised
    from real_robots_b200.policy import BatchedPolicy
    ve.VecREALRobotEnv = _FakeVecEnv

    class Push(BatchedPolicy):
        first = None

        def step(self, observation, reward, done):
            if Push.first is None:        # evaluate.py:285: the first observation already shows the goal's initial placement
                Push.first = observation["object_positions"].clone()
                assert observation["rendered"] is False
            return {"shift": torch.tensor([[0.001, 0.0, 0.0]], dtype=torch.float64).expand(self.env.n_envs, 3)}
    challenge = ['2D' if g % 2 == 0 else '3D' for g in range(n_goals)]
    so, scores, vec = evaluate_vectorised(Push, _goals(n_goals), n_objects=2, action_type='joints', extrinsic_timesteps=7,
                                          challenge_of=challenge)
    lo, hi = shard_range(n_goals, rank, world)
    if hi > lo:
        assert np.allclose(Push.first[:hi - lo].numpy(), _goals(n_goals)["initial"][lo:hi, :, :3])
    q.put((rank, vec, scores, {k: float(v) for k, v in so.items()}, _FakeVecEnv.made[-1].n_envs))
    if world > 1:
        dist.destroy_process_group()


def _run_eval(world, n_goals):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_eval_worker, args=(r, world, port, n_goals, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = {r: rest for r, *rest in (q.get(timeout=180) for _ in range(world))}
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return res


def test_evaluate_vectorised_shards_goals_gloo_world2():
    """Trial-parallel extrinsic phase across two ranks (ragged 3 + 2 goals, then 1 goal = an empty shard on rank 1):
    every rank ends with the same dense score vector, equal to the single-process result."""
    for n_goals in (5, 1):
        goals = _goals(n_goals)
        expect = _expected_scores(goals, 7)
        one = _run_eval(1, n_goals)[0]
        two = _run_eval(2, n_goals)
        assert np.allclose(one[0], expect, atol=1e-12)
        for r in (0, 1):
            vec, scores, so, n_env = two[r]
            assert np.allclose(vec, expect, atol=1e-12)
            assert scores == one[1] and so == one[2]
            lo, hi = shard_range(n_goals, r, 2)
            assert n_env == max(hi - lo, 1)
        assert sorted(one[1].keys()) == (['2D', '3D'] if n_goals > 1 else ['2D'])
