"""Distribution of per-env solver work (rows, contacts) in the steady state of scripts/time_step.py's scenario and of bench.py's."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch, numpy as np
from real_robots_b200.sim import BatchedSim
import bench
N = 4096
for name in ("time_step", "bench"):
    sim = BatchedSim(N, 3)
    if name == "time_step":
        g = torch.Generator(device="cuda"); g.manual_seed(0)
        acts = [(torch.rand(N, 9, dtype=torch.float64, device="cuda", generator=g) * 2 - 1) * 2.0 for _ in range(8)]
        for i in range(120): sim.step(acts[(i // 20) % 8])
    else:
        acts = bench.random_policy_actions(torch, 200, N, torch.device("cuda"), 1234)
        for i in range(200): sim.step(acts[i])
    d = sim.diag().cpu().numpy()
    rows, nc, fl = d[:, 0], d[:, 2], d[:, 3]
    print(name, "rows pct 50/90/99/max", np.percentile(rows, [50, 90, 99]).tolist(), rows.max(), "contacts 50/90/99/max", np.percentile(nc, [50, 90, 99]).tolist(), nc.max(),
          "non-simple frac %.3f" % ((fl & 2) == 0).mean(), "cap hit", int((fl & 1).sum()))
    cost = (fl >> 16) & 0xffff
    print("   sweep cost (slots/iteration, pair slots twice): pct 50/90/96/98/99/99.9", np.percentile(cost, [50, 90, 96, 98, 99, 99.9]).tolist(), "max", cost.max(),
          "classes >72 / >48 / >32 / rest:", int((cost > 72).sum()), int(((cost > 48) & (cost <= 72)).sum()), int(((cost > 32) & (cost <= 48)).sum()), int((cost <= 32).sum()),
          " sum of cost %d, of which cost>32: %d" % (cost.sum(), cost[cost > 32].sum()))
    print("   histogram (cost: envs):", {int(c): int((cost == c).sum()) for c in np.unique(cost)})
    sim.close()
