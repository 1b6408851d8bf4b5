"""Where a C5 step (cartesian action, R1 observation) spends its time: kernel event times from the library's profiler, IK timed apart."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from real_robots_b200.envs.vec_env import VecREALRobotEnv
n = int(os.environ.get("RR_N", "4096"))
dev = torch.device("cuda")
env = VecREALRobotEnv(n, objects=3, action_type='cartesian', additional_obs=True)
env.reset(render=False)
g = torch.Generator(device=dev); g.manual_seed(5)
lo = torch.tensor([-0.25, -0.5, 0.40, -1, -1, -1, -1], dtype=torch.float64, device=dev)
hi = torch.tensor([0.25, 0.5, 0.60, 1, 1, 1, 1], dtype=torch.float64, device=dev)
cart = lo + (hi - lo) * torch.rand(n, 7, dtype=torch.float64, device=dev, generator=g)
grip = torch.rand(n, 2, dtype=torch.float64, device=dev, generator=g) * 1.5707963
def step(t):
    if t % 20 == 0:
        cart.copy_(lo + (hi - lo) * torch.rand(n, 7, dtype=torch.float64, device=dev, generator=g))
    return env.step({'cartesian_command': cart, 'gripper_command': grip, 'render': True})
for t in range(60): step(t)
torch.cuda.synchronize()
env.sim.profile(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for t in range(40): step(t)
e1.record(); torch.cuda.synchronize()
prof = env.sim.profile_read(); env.sim.profile(False)
print("C5 %d envs: %.3f ms per step; kernel event ms per step:" % (n, e0.elapsed_time(e1) / 40), {k: round(v[0] / 40, 3) for k, v in prof.items()})
d = env.sim.diag(); cost = (d[:, 3] >> 16) & 0xffff
print("   schedule length: median %d, 90 %% %d, max %d, envs above 24: %d" % (int(cost.median()), int(cost.double().quantile(0.9)), int(cost.max()), int((cost > 24).sum())))
# IK alone
e0.record()
for i in range(3):
    c = lo + (hi - lo) * torch.rand(n, 7, dtype=torch.float64, device=dev, generator=g)
    env.sim.ik(c.contiguous(), 1000, 0.001)
e1.record(); torch.cuda.synchronize()
print("   rr_ik (1000 iterations max, residual 1e-3) on fresh random targets: %.2f ms per call" % (e0.elapsed_time(e1) / 3))
