"""Times rr_render (and checks a checksum) for the library selected with RR_LIB."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from real_robots_b200.sim import BatchedSim
N = 4096
sim = BatchedSim(N, 3)
g = torch.Generator(device="cuda"); g.manual_seed(0)
act = (torch.rand(N, 9, dtype=torch.float64, device="cuda", generator=g) * 2 - 1) * 2.0
for i in range(60): sim.step(act)
for _ in range(3): sim.render()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): rgb, depth, seg = sim.render()
e1.record(); torch.cuda.synchronize()
chk = int(rgb.to(torch.int64).sum()) ^ int(seg.to(torch.int64).sum()) ^ int((depth.double() * 1e6).sum())
print("%s: render %.3f ms  checksum %d" % (os.path.basename(os.environ.get("RR_LIB", "default")), e0.elapsed_time(e1) / 10, chk))
