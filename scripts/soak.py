"""Soak of the overlapped schedule: 4096 envs, random policy, step + render + read_state for many steps; no time-out of the in-kernel
wait, no fault flags, finite state, and the final state equals that of a context stepped with the sequential schedule."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import bench
from real_robots_b200.sim import BatchedSim
N, T = 4096, int(os.environ.get("SOAK_STEPS", "1500"))
acts = bench.random_policy_actions(torch, T, N, torch.device("cuda"), 99)
sim = BatchedSim(N, 3)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
worst_heavy = 0
for t in range(T):
    sim.step(acts[t]); sim.render(); sim.read_state()
    if t % 100 == 99:
        d = sim.diag()
        worst_heavy = max(worst_heavy, int((((d[:, 3] >> 16) & 0xffff) > 24).sum()))
e1.record(); torch.cuda.synchronize()
q, qd, obj = sim.get_state()
print("soak: %d steps, %.3f ms per step, finite %s, faults %d, most heavy envs in a step %d" % (
    T, e0.elapsed_time(e1) / T, bool(torch.isfinite(q).all() and torch.isfinite(obj).all()), int((sim.faults() != 0).sum()), worst_heavy))
os.environ["RR_HEAVY_SOLVER"] = "0"; os.environ["RR_EAGER_COLLIDE"] = "0"
ref = BatchedSim(N, 3)
for t in range(T):
    ref.step(acts[t])
same = all(torch.equal(a, b) for a, b in zip(sim.get_state(), ref.get_state()))
print("soak: final state equals the sequential schedule's bit for bit:", same)
assert same
