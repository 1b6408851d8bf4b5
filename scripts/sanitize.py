"""Small run for compute-sanitizer: physics with contacts, IK, render, host path."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from real_robots_b200.sim import BatchedSim
N = 6
sim = BatchedSim(N, 3)
rng = np.random.default_rng(0)
lo = np.array([-2.09, -2.09, -2.96, -2.09, -2.96, -2.09, -3.05, 0, 0]); hi = -lo; hi[7:] = np.pi / 2
a = torch.from_numpy(rng.uniform(lo, hi, (N, 9))).cuda()
for t in range(45):
    if t == 30: a = torch.from_numpy(rng.uniform(lo, hi, (N, 9))).cuda()
    sim.step(a)
for t in range(6):                 # render straight after the step: k_raster's late-env path beside the heavy-env solver kernel and the look-ahead narrowphase
    sim.step(a); sim.render(); sim.read_state()
sim.set_object_pose(0, torch.tensor([[-0.1, 0.0, 0.6, 0, 0, 0, 1.0]] * N, dtype=torch.float64, device="cuda"))   # drops the look-ahead narrowphase
sim.step(a)
sim.read_state(); sim.render()
sim.render(mask=torch.tensor([1, 0, 1, 0, 1, 1], dtype=torch.uint8, device="cuda"))
pq = torch.tensor([[-0.1, 0.0, 0.5, 0.7, 0.7, 0, 0]] * N, dtype=torch.float64, device="cuda")
sim.ik(pq, 200, 1e-3)
sim.step_host(a.cpu().numpy(), np.ones(N, np.uint8))
sim.render_frames(sim.shape_frames())                       # every shape rasterised (k_raster<0>); rr_render uses the static layer (k_raster<2>)
from real_robots_b200.cameras import view_from_yaw_pitch_roll, projection_fov
sim.render_camera(view_from_yaw_pitch_roll([0, 0, 0.4], 1.2, 30, -30, 0), projection_fov(80, 320 / 240, 0.1, 100.0))   # other-camera path
sim.contacts(); sim.shape_frames(); sim.link_pose(7); sim.reset()
torch.cuda.synchronize()
print("sanitize run ok")
