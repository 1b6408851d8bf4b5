"""Physics-only timing under the bench's random policy (arm contacts included) for the library selected with RR_LIB."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import bench
from real_robots_b200.sim import BatchedSim
N = int(os.environ.get("RR_N", "4096"))
sim = BatchedSim(N, 3)
acts = bench.random_policy_actions(torch, 160, N, torch.device("cuda"), 1234)
for t in range(100): sim.step(acts[t])
torch.cuda.synchronize()
sim.profile(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for t in range(100, 160): sim.step(acts[t])
e1.record(); torch.cuda.synchronize()
wall = e0.elapsed_time(e1) / 60
prof = sim.profile_read()
q = sim.read_state()
chk = float(q[0].sum()) + float(q[2].sum())
dg = sim.diag(); d = dg.double().mean(dim=0).tolist(); d[3] = float(((dg[:, 3] & 2) != 0).double().mean())
print("%s: step %.3f ms (physics only, back to back); dyn+solve %.3f ms, collide %.3f ms  checksum %.12f  diag(rows, iters, contacts, overflow) %s" % (
    os.path.basename(os.environ.get("RR_LIB", "default")), wall, prof["step"][0] / max(prof["step"][1], 1),
    prof["collide"][0] / max(prof["collide"][1], 1), chk, ["%.2f" % x for x in d]))
