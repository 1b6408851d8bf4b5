"""Wall time per step (CUDA events, no profiler) of three call patterns at 4096 envs under the bench's random policy, for the schedule
switches in the environment: step only; step + read_state; step + render + read_state."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import bench
from real_robots_b200.sim import BatchedSim
N = 4096
acts = bench.random_policy_actions(torch, 200, N, torch.device("cuda"), 1234)
def run(pattern, steps=80):
    sim = BatchedSim(N, 3)
    for t in range(100): sim.step(acts[t])
    def one(t):
        sim.step(acts[100 + t % 100])
        if pattern >= 2: sim.render()
        if pattern >= 1: sim.read_state()
    for t in range(10): one(t)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(10, 10 + steps): one(t)
    e1.record(); torch.cuda.synchronize()
    sim.close()
    return e0.elapsed_time(e1) / steps
print("RR_HEAVY_SOLVER=%s RR_EAGER_COLLIDE=%s: step %.3f | step+read %.3f | step+render+read %.3f ms" % (
    os.environ.get("RR_HEAVY_SOLVER", "-"), os.environ.get("RR_EAGER_COLLIDE", "-"), run(0), run(1), run(2)))
