"""Times rr_step (collide + dynamics + solve) in the steady state for the library selected with RR_LIB."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from real_robots_b200.sim import BatchedSim
N = 4096
sim = BatchedSim(N, 3)
g = torch.Generator(device="cuda"); g.manual_seed(0)
acts = [(torch.rand(N, 9, dtype=torch.float64, device="cuda", generator=g) * 2 - 1) * float(os.environ.get("ACT_SCALE", "2.0")) for _ in range(8)]
for i in range(80): sim.step(acts[(i // 20) % 8])
torch.cuda.synchronize()
sim.profile(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(40): sim.step(acts[(i // 20 + 4) % 8])
e1.record(); torch.cuda.synchronize()
prof = sim.profile_read()
q = sim.read_state()
chk = float(q[0].sum()) + float(q[2].sum())
d = sim.diag().double().mean(dim=0).tolist() if hasattr(sim, "diag") else []
print("%s: step %.3f ms  (dyn+solve %.3f, collide %.3f)  checksum %.12f  diag(rows, iters, contacts, overflow) %s" % (
    os.path.basename(os.environ.get("RR_LIB", "default")), e0.elapsed_time(e1) / 40, prof["step"][0] / max(prof["step"][1], 1),
    prof["collide"][0] / max(prof["collide"][1], 1), chk, ["%.2f" % x for x in d]))
