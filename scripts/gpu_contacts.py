"""Dev check: contact-rich parity CUDA vs oracle."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from real_robots_b200.sim import BatchedSim
from oracle.orc import Oracle
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
rng = np.random.default_rng(3)
sim = BatchedSim(N, 3)
orcs = [Oracle(3) for _ in range(N)]
lo = np.array([-2.09, -2.09, -2.96, -2.09, -2.96, -2.09, -3.05, 0, 0]); hi = -lo; hi[7:] = np.pi / 2
act = np.zeros((N, 9))
first_div = [None] * N
for t in range(steps):
    if t % 40 == 0 and t > 0:
        act = rng.uniform(lo, hi, size=(N, 9)); act[0] = 0
    sim.step(torch.from_numpy(act).cuda())
    q, qd, obj = [x.cpu().numpy() for x in sim.get_state()]
    d = sim.diag().cpu().numpy()
    for e in range(N): orcs[e].step(act[e])
    oq = np.array([o.get_state()[0] for o in orcs]); oobj = np.array([o.get_state()[2] for o in orcs])
    od = [o.diag() for o in orcs]
    eq = abs(q - oq).max(axis=1); eo = abs(obj - oobj).max(axis=(1, 2))
    for e in range(N):
        if first_div[e] is None and (eq[e] > 1e-6 or eo[e] > 1e-6): first_div[e] = t
    if t % 20 == 0 or t == steps - 1:
        print("t %3d |dq| %.2e |dobj| %.2e  contacts gpu %s orc %s  iters gpu %s orc %s ovf %d" % (
            t, eq.max(), eo.max(), d[:4, 2], [x['contacts'] for x in od[:4]], d[:4, 1], [x['iterations'] for x in od[:4]], d[:, 3].sum()))
print("first divergence step per env (>1e-6):", first_div)
j, tch, op = sim.read_state()
print("touch gpu", tch[:4].cpu().numpy().round(3)); print("touch orc", np.array([o.read()[1] for o in orcs[:4]]).round(3))
print("obj z gpu", obj[:4, :, 2].round(4)); print("obj z orc", oobj[:4, :, 2].round(4))
