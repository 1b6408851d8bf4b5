import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from real_robots_b200.sim import BatchedSim
sim = BatchedSim(1, 3)
a = torch.zeros(1, 9, dtype=torch.float64, device="cuda"); a[:, 1] = 0.5; a[:, 3] = -0.8
for t in range(40): sim.step(a)
sim.read_state()
for t in range(4): sim.render()
torch.cuda.synchronize()
