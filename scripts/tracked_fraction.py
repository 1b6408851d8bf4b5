"""CUDA path vs oracle in contact-rich random-policy episodes: tracked fraction (free running) and one-step agreement from
identical states, for the library selected with RR_LIB (profiles/r02_divergence.md compares the default build with a
-fmad=false build of rr_physics.cu)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from real_robots_b200.sim import BatchedSim
from tests.scenarios import contact_rich_tracking
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
for seed in (3, 4):
    N = 32
    sim = BatchedSim(N, 3)
    r = contact_rich_tracking(lambda a: sim.step(dev(a)), lambda: [t.cpu().numpy() for t in sim.get_state()],
                              lambda q, qd, obj: sim.set_state(dev(q), dev(qd), dev(obj)), N=N, T=400, seed=seed)
    print("%s seed %d: tracked (400 free-running steps, 1e-6) %.3f   one-step agreement (1e-9) %.4f   largest flip %.2e" % (
        os.path.basename(os.environ.get("RR_LIB", "default")), seed, *r))
    sim.close()
