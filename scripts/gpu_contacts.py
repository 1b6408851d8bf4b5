"""Dev check: contact-rich parity CUDA vs oracle under the random policy (cli.py:12-20 semantics)."""
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
act = rng.uniform(lo, hi, size=(N, 9))
first_div = [None] * N
arm_steps = np.zeros(N, int); touch_steps = np.zeros(N, int)
for t in range(steps):
    res = rng.random(N) < 0.05
    new = rng.uniform(lo, hi, size=(N, 9))
    act = np.where(res[:, None], new, act)
    sim.step(torch.from_numpy(act).cuda())
    q, qd, obj = [x.cpu().numpy() for x in sim.get_state()]
    d = sim.diag().cpu().numpy()
    tch = sim.read_state()[1].cpu().numpy()
    for e in range(N): orcs[e].step(act[e])
    oq = np.array([o.get_state()[0] for o in orcs]); oobj = np.array([o.get_state()[2] for o in orcs])
    otch = np.array([o.read()[1] for o in orcs])
    eq = abs(q - oq).max(axis=1); eo = abs(obj - oobj).max(axis=(1, 2))
    for e in range(N):
        A, B, D = orcs[e].contacts()
        if (A < 17).any(): arm_steps[e] += 1
        if otch[e].max() > 0: touch_steps[e] += 1
        if first_div[e] is None and (eq[e] > 1e-6 or eo[e] > 1e-6): first_div[e] = t
    if t % 50 == 0 or t == steps - 1:
        ok = np.array([f is None for f in first_div])
        print("t %3d  envs still within 1e-6: %d/%d  max|dq| among them %.2e  |dobj| %.2e  touch err %.2e  contacts(max) gpu %d  ovf %d  finite %s" % (
            t, ok.sum(), N, eq[ok].max() if ok.any() else -1, eo[ok].max() if ok.any() else -1, abs(tch - otch)[ok].max() if ok.any() else -1,
            d[:, 2].max(), d[:, 3].sum(), np.isfinite(q).all() and np.isfinite(obj).all()))
print("first divergence step per env (>1e-6):", first_div)
print("oracle steps with arm contact per env:", arm_steps.tolist())
print("oracle steps with touch>0 per env:", touch_steps.tolist())
# statistical comparison at the end (all envs): object heights and joint positions
print("final obj z gpu mean %s  orc mean %s" % (obj[:, :, 2].mean(0).round(4), oobj[:, :, 2].mean(0).round(4)))
print("final |q| gpu mean %.4f orc mean %.4f" % (abs(q).mean(), abs(oq).mean()))
