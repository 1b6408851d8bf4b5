"""BASELINE configs[4]: 65 536 envs across 8 B200 (8192 per GPU), cartesian action (xyz ~ U(env.py:45-46), quaternion components
~ U(-1, 1) un-normalised, gripper ~ U(0, pi/2)), full R1 observation dict with the retina, through the public
batched API (VecREALRobotEnv), one rank per GPU under torchrun, no collective on the path.  A new cartesian request (= a new
IK solve for every env) every 20th step.  Prints one JSON line on rank 0: aggregate env-steps/s, max over ranks.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 scripts/bench_c5.py [--envs 8192]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import torch.distributed as dist
from real_robots_b200.envs.vec_env import VecREALRobotEnv

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=8192)
ap.add_argument("--steps", type=int, default=40)
a = ap.parse_args()
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n = a.envs
env = VecREALRobotEnv(n, device=local, objects=3, action_type='cartesian', additional_obs=True)
env.reset(render=False)
g = torch.Generator(device=dev); g.manual_seed(5 + rank)
lo = torch.tensor([-0.25, -0.5, 0.40, -1, -1, -1, -1], dtype=torch.float64, device=dev)
hi = torch.tensor([0.25, 0.5, 0.60, 1, 1, 1, 1], dtype=torch.float64, device=dev)
cart = lo + (hi - lo) * torch.rand(n, 7, dtype=torch.float64, device=dev, generator=g)
grip = torch.rand(n, 2, dtype=torch.float64, device=dev, generator=g) * 1.5707963


def step(t):
    if t % 20 == 0:
        cart.copy_(lo + (hi - lo) * torch.rand(n, 7, dtype=torch.float64, device=dev, generator=g))
    return env.step({'cartesian_command': cart, 'gripper_command': grip, 'render': True})


for t in range(60):
    obs, _, _, _ = step(t)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for t in range(a.steps):
    obs, _, _, _ = step(t)
e1.record(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    keys = sorted(obs.keys())
    print(json.dumps({"config": "C5: cartesian action, full R1 observation, %d envs/GPU x %d GPU = %d envs" % (n, world, n * world),
                      "env_steps_per_s": world * n * a.steps / (float(ms.item()) * 1e-3), "ms_per_step": float(ms.item()) / a.steps,
                      "n_gpus": world, "envs_per_gpu": n, "steps": a.steps, "obs_keys": keys}))
if world > 1:
    dist.destroy_process_group()
