"""Development check run under gpurun: CUDA path vs oracle, prints error magnitudes (not a test)."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from real_robots_b200.sim import BatchedSim
from oracle.orc import Oracle

def physics(N=32, steps=60, contacts=0, seed=0):
    rng = np.random.default_rng(seed)
    sim = BatchedSim(N, 3, enable_contacts=contacts)
    orcs = [Oracle(3, enable_contacts=contacts) for _ in range(N)]
    lo = np.array([-2.09, -2.09, -2.96, -2.09, -2.96, -2.09, -3.05, 0, 0]); hi = -lo; hi[7:] = np.pi / 2
    act = rng.uniform(lo, hi, size=(N, 9))
    worst = 0
    for t in range(steps):
        if t % 20 == 0:
            act = rng.uniform(lo, hi, size=(N, 9))
        sim.step(torch.from_numpy(act).cuda())
        q, qd, obj = [x.cpu().numpy() for x in sim.get_state()]
        for e in range(N):
            orcs[e].step(act[e])
        oq = np.array([o.get_state()[0] for o in orcs]); oqd = np.array([o.get_state()[1] for o in orcs]); oobj = np.array([o.get_state()[2] for o in orcs])
        eq, eqd, eo = abs(q - oq).max(), abs(qd - oqd).max(), abs(obj - oobj).max()
        worst = max(worst, eq)
        if t % 10 == 0 or t == steps - 1:
            d = sim.diag().cpu().numpy()
            print("step %3d  |dq| %.3e  |dqd| %.3e  |dobj| %.3e  iters gpu %s oracle %s" % (t, eq, eqd, eo, d[:3, 1], [o.diag()['iterations'] for o in orcs[:3]]))
    j, tch, op = sim.read_state()
    print("joint_pos[0]", j[0].cpu().numpy()); print("oracle     ", orcs[0].read()[0])
    return worst

def render(N=8, seed=1):
    rng = np.random.default_rng(seed)
    sim = BatchedSim(N, 3, enable_contacts=0)
    orc = Oracle(3, enable_contacts=0)
    frames = []
    refs = []
    for e in range(N):
        q = rng.uniform(-1.5, 1.5, 11) if e else np.zeros(11)
        q[7:] = rng.uniform(0, 1.0, 4) * (e > 0)
        obj = np.zeros((3, 13))
        for o in range(3):
            obj[o, :3] = [rng.uniform(-0.25, 0.05), rng.uniform(-0.4, 0.4), rng.uniform(0.32, 0.5)]
            qq = rng.normal(size=4); obj[o, 3:7] = qq / np.linalg.norm(qq)
        if e == 0:
            orc.reset()
        else:
            orc.set_state(q=q, qd=np.zeros(11), obj=obj)
        frames.append(orc.shape_frames())
        refs.append(orc.render())
    frames = np.stack(frames)
    t0 = time.time()
    rgb, depth, seg = sim.render_frames(torch.from_numpy(frames).cuda())
    torch.cuda.synchronize()
    rgb, depth, seg = rgb.cpu().numpy(), depth.cpu().numpy(), seg.cpu().numpy()
    for e in range(N):
        r, d, m = refs[e]
        print("env %d: mask mismatches %d  rgb mismatches %d (max %d)  depth mismatches %d (max %.3e)  nonbg %d" % (
            e, (m != seg[e]).sum(), (r != rgb[e]).any(-1).sum(), abs(r.astype(int) - rgb[e].astype(int)).max(),
            (d != depth[e]).sum(), abs(d - depth[e]).max(), (m >= 0).sum()))
    os.makedirs("gpurun_out", exist_ok=True)
    np.savez_compressed("gpurun_out/render_check.npz", rgb=rgb[:4], ref=np.stack([r[0] for r in refs[:4]]), seg=seg[:4], refseg=np.stack([r[2] for r in refs[:4]]))

def speed(N=4096, steps=20):
    sim = BatchedSim(N, 3, enable_contacts=0)
    act = torch.zeros(N, 9, dtype=torch.float64, device="cuda").uniform_(-1, 1)
    for name, fn in (("step", lambda: sim.step(act)), ("render", lambda: sim.render()), ("step+render", lambda: (sim.step(act), sim.render()))):
        for _ in range(3): fn()
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        print("%-12s N=%d: %.3f ms/step  -> %.0f env-steps/s" % (name, N, ms, N / ms * 1e3))

if __name__ == "__main__":
    which = sys.argv[1:] or ["physics", "render", "speed"]
    print(torch.cuda.get_device_name(0))
    if "physics" in which: physics()
    if "render" in which: render()
    if "speed" in which: speed()
