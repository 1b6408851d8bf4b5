"""rr_render for small batches (strips of an env spread over several CTAs), ms per launch."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from real_robots_b200.sim import BatchedSim
for n in (1, 2, 8, 48, 49, 98, 148, 296):
    sim = BatchedSim(n, 3)
    a = torch.zeros(n, 9, dtype=torch.float64, device="cuda"); a[:, 1] = 0.5; a[:, 3] = -0.8
    for t in range(40): sim.step(a)
    sim.read_state()
    for t in range(5): sim.render()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(50): sim.render()
    e1.record(); torch.cuda.synchronize()
    print("%4d envs: rr_render %.4f ms" % (n, e0.elapsed_time(e1) / 50))
    sim.close()
