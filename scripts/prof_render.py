"""Profiling driver: a few steps + renders at N envs (run under ncu via gpurun)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from real_robots_b200.sim import BatchedSim
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
contacts = int(sys.argv[2]) if len(sys.argv) > 2 else 0
sim = BatchedSim(N, 3, enable_contacts=contacts)
g = torch.Generator(device="cuda"); g.manual_seed(0)
act = (torch.rand(N, 9, dtype=torch.float64, device="cuda", generator=g) * 2 - 1) * 2.0
for i in range(40):
    sim.step(act)
for i in range(3):
    sim.step(act); sim.render()
torch.cuda.synchronize()
