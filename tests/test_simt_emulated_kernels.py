"""CPU tests that execute the SHIPPED CUDA kernel sources (real_robots_b200/csrc/*.cu) under the SIMT emulator of
tests/simt (every CUDA thread a fiber; shuffles, ballots, barriers, shared memory and atomics emulated) and hold them to
the oracle: dynamics + narrowphase + solver to 1e-9 per step (the emulated build contracts no FMAs, so it is in fact
~1e-15), the rasteriser bit for bit.  The GPU parity tests (-m gpu) repeat this on hardware; these run without a GPU so that
a kernel change is checked before GPU time is spent.  Test infrastructure only — the package cannot load the emulated build."""
import os
import numpy as np
import pytest

from oracle.orc import Oracle
from tests.simt.emu import EmuSim

LO = np.array([-2.09, -2.09, -2.96, -2.09, -2.96, -2.09, -3.05, 0, 0])
HI = -LO
HI[7:] = np.pi / 2


def _policy(rng, n, steps):
    a = rng.uniform(LO, HI, size=(n, 9))
    for _ in range(steps):
        redraw = rng.random(n) < 0.05
        a = np.where(redraw[:, None], rng.uniform(LO, HI, size=(n, 9)), a)
        yield a.copy()


def _policy_single(seed, steps):
    rng = np.random.default_rng(seed)
    a = rng.uniform(LO, HI)
    for _ in range(steps):
        if rng.random() < 0.05:
            a = rng.uniform(LO, HI)
        yield a.copy()


def test_emulated_step_kernels_match_oracle_with_contacts():
    """k_collide + k_step_dyn + k_step_solve on 3 envs x 150 random-policy steps (objects land at ~step 27, the arm hits the
    table / objects later): joints, touch, object poses and solver diagnostics against the oracle every step, to 1e-9
    (north_star: 1e-6 per substep).  Contact dynamics amplify rounding differences (the kernel sums rows in a tree, the
    oracle sequentially) by ~1.2x per step, so every 15 steps both sides restart from the oracle's state (which also drops
    both manifold caches, so cold and warm-started contacts are both exercised)."""
    N = 3
    seeds = [11, 13, 19]                         # 13 and 19: the arm lies on objects / the table for dozens of steps (rows coupling arm and object)
    sim = EmuSim(N, 3)
    orcs = [Oracle(3) for _ in range(N)]
    seen_contacts, seen_wide = 0, 0
    for s, a in enumerate(zip(*[_policy_single(sd, 150) for sd in seeds])):
        a = np.stack(a)
        if s % 15 == 0 and s > 0:
            st = [x.get_state() for x in orcs]
            for e in range(N):
                orcs[e].set_state(*st[e])
            sim.set_state(np.stack([x[0] for x in st]), np.stack([x[1] for x in st]), np.stack([x[2] for x in st]))
        sim.step(a)
        for e in range(N):
            orcs[e].step(a[e])
        j, t, o = sim.read_state()
        d = sim.diag()
        for e in range(N):
            rj, rt = orcs[e].read()
            od = orcs[e].diag()
            assert np.abs(j[e] - rj).max() < 1e-9, (s, e)
            assert np.abs(t[e] - rt).max() < 1e-6 * max(1.0, np.abs(rt).max()), (s, e)
            assert d[e, 2] == od["contacts"] and d[e, 1] == od["iterations"] and d[e, 0] == od["rows"], (s, e, d[e], od)
            assert (d[e, 3] & 0xffff & ~2) == 0                 # no cap, no fault, no dropped manifold
            seen_contacts = max(seen_contacts, od["contacts"])
            seen_wide += int(od["contacts"] > 0 and not (d[e, 3] & 2))
        q, qd, ob = sim.get_state()
        ref = np.array([x.get_state()[2] for x in orcs])
        assert np.abs(ob - ref).max() < 1e-9, s
    assert seen_contacts >= 8 and seen_wide >= 20, (seen_contacts, seen_wide)        # arm contacts (rows coupling the arm with an object) were solved too


def test_emulated_rasteriser_is_bit_exact_and_static_layer_is_identical():
    N = 2
    rng = np.random.default_rng(5)
    sim = EmuSim(N, 3)
    orcs = [Oracle(3) for _ in range(N)]
    for a in _policy(rng, N, 40):
        sim.step(a)
        for e in range(N):
            orcs[e].step(a[e])
    frames = np.stack([o.shape_frames() for o in orcs])
    rgb, depth, seg = sim.render_frames(frames)                    # k_raster<0>: every shape rasterised
    for e in range(N):
        r, d, m = orcs[e].render()
        assert (m == seg[e]).all() and (r == rgb[e]).all() and (d == depth[e]).all()
    rgb2, depth2, seg2 = sim.render()                              # k_raster<2>: on top of the static layer built by k_raster<1>
    rgb3, depth3, seg3 = sim.render_frames(sim.shape_frames())
    assert (rgb2 == rgb3).all() and (depth2 == depth3).all() and (seg2 == seg3).all()
    assert (seg2 >= 0).sum() > 1000 and set(np.unique(seg2)) >= {-1, 0, 1, 2}


def test_emulated_masked_render_returns_zero_images():
    sim = EmuSim(2, 1)
    rgb, depth, seg = sim.render(mask=np.array([0, 1], np.uint8))
    assert not rgb[0].any() and not depth[0].any() and not seg[0].any()
    assert rgb[1].any() and (seg[1] >= 0).any()


def test_emulated_ik_matches_oracle():
    rng = np.random.default_rng(2)
    N = 4
    sim = EmuSim(N, 1, enable_contacts=0)
    o = Oracle(1, enable_contacts=0)
    pq = np.zeros((N, 7))
    pq[:, :3] = rng.uniform([-0.3, -0.3, 0.45], [0.0, 0.3, 0.7], size=(N, 3))
    pq[:, 3:] = [1.0, 0.0, 0.0, 0.0]
    out = sim.ik(pq)
    for e in range(N):
        ref, _ = o.ik(pq[e, :3], pq[e, 3:])
        assert np.abs(out[e] - ref).max() < 1e-8


def test_contact_free_aabb_overlaps_do_not_crowd_out_object_table_manifolds():
    """manifold slots used to be given to every AABB-overlapping pair, contact or not, so that with > 24
    overlaps the object x table pairs (last in the pair order) lost their manifolds and the objects fell through the table
    under the reference's own RandomPolicy.  Now only non-empty manifolds take a slot (and a dropped one is flagged).
    Oracle: 6000 random-policy steps with several steps of > 24 overlaps, no object ever below the table top inside the
    table's footprint.  Emulated CUDA kernels: replay of the first crowded stretch from the oracle's state: no flag, objects stay on the table."""
    rng = np.random.default_rng(41)
    lo = np.ones(9) * -np.pi * 0.944; hi = np.ones(9) * np.pi * 0.944
    lo[0] = -np.pi * 0.666; hi[0] = np.pi * 0.666; lo[1:9:2] = -np.pi * 0.666; hi[1:9:2] = np.pi * 0.666
    lo[6] = -np.pi * 0.972; hi[6] = np.pi * 0.972; lo[-2:] = 0; hi[-2:] = np.pi / 2
    o = Oracle(3)
    a = rng.uniform(lo, hi)
    crowded, min_z, acts, states, first = 0, 9.0, [], {}, None
    for t in range(6000):
        if rng.random() < 0.05:
            a = rng.uniform(lo, hi)
        if t % 20 == 0:
            states[t] = o.get_state()
        acts.append(a.copy())
        o.step(a)
        d = o.diag2()
        assert d["manifold_overflow"] == 0
        if d["overlaps"] > 24:
            crowded += 1
            first = t if first is None else first
        ob = o.get_state()[2]
        inside = (ob[:, 0] > -0.22) & (ob[:, 0] < 0.02) & (np.abs(ob[:, 1]) < 0.42)
        if inside.any():
            min_z = min(min_z, ob[inside, 2].min())
    assert crowded >= 10 and first is not None
    assert min_z > 0.29, min_z                   # table top 0.2794 + the flattest object's half height
    t0 = (first // 20) * 20
    q, qd, obj = states[t0]
    ref = Oracle(3); ref.set_state(q, qd, obj)
    sim = EmuSim(1, 3); sim.set_state(q[None], qd[None], obj[None])
    seen = 0
    for t in range(t0, min(t0 + 60, 6000)):
        ref.step(acts[t]); sim.step(acts[t][None])
        seen = max(seen, ref.diag2()["overlaps"])
        # face-to-face penetrations make the EPA witness point jump between coplanar faces on 1e-15 pose differences, so a
        # contact-rich replay is not held to 1e-9 (test_emulated_step_kernels... does that where contacts are well conditioned)
        assert np.abs(sim.read_state()[0][0, :7] - ref.read()[0][:7]).max() < 0.1      # arm joints stay together; the fingers chatter on the table
        ob = sim.get_state()[2][0]
        inside = (ob[:, 0] > -0.22) & (ob[:, 0] < 0.02) & (np.abs(ob[:, 1]) < 0.42)
        assert not inside.any() or ob[inside, 2].min() > 0.29
        assert (sim.diag()[0, 3] & 16) == 0
    assert seen > 24


def test_emulated_host_buffer_path_and_pipelined_begin_end():
    """rr_step_host == rr_step + rr_read_state + rr_render; rr_step_host_begin / _end with two steps in flight gives the same
    observations (double-buffered staging), and a third begin without an end is refused."""
    N = 2
    a = np.stack(list(zip(*[_policy_single(sd, 6) for sd in (3, 4)])))          # [6, N, 9]
    ref = EmuSim(N, 3)
    want = []
    for t in range(6):
        ref.step(a[t])
        j, tch, o = ref.read_state()
        rgb, depth, seg = ref.render()
        want.append((j.copy(), tch.copy(), o.copy(), rgb.copy(), depth.copy(), seg.copy()))
    sim = EmuSim(N, 3)
    for t in range(2):
        out = sim.step_host(a[t])
        for got, w in zip((out["joint_positions"], out["touch_sensors"], out["object_poses"], out["retina"], out["depth"], out["mask"]), want[t]):
            assert (got == w).all()
    mask = np.ones(N, np.uint8)
    outs = [sim.alloc_host(), sim.alloc_host()]
    cmds = [np.ascontiguousarray(a[t]) for t in range(6)]
    sim.step_host_begin(cmds[2], mask, outs[0])
    sim.step_host_begin(cmds[3], mask, outs[1])
    with pytest.raises(RuntimeError):
        sim.step_host_begin(cmds[4], mask, outs[0])
    sim.step_host_end()
    assert (outs[0]["retina"] == want[2][3]).all() and (outs[0]["joint_positions"] == want[2][0]).all()
    sim.step_host_begin(cmds[4], mask, outs[0])
    sim.step_host_end()
    assert (outs[1]["mask"] == want[3][5]).all() and (outs[1]["depth"] == want[3][4]).all()
    sim.step_host_end()
    assert (outs[0]["retina"] == want[4][3]).all() and (outs[0]["object_poses"] == want[4][2]).all()
    with pytest.raises(RuntimeError):
        sim.step_host_end()


def test_emulated_touch_sensors_scripted_press_and_shelf():
    from tests.scenarios import check_touch_scenarios
    check_touch_scenarios(lambda: EmuSim(1, 1))


def test_integration_md_binding_stub_runs_against_the_c_abi():
    """INTEGRATION.md section B's B200Client (the binding a maintainer of the reference would add), executed verbatim —
    here against the emulated build of the library; tests/test_gpu_facade.py does the same against libreal_b200.so."""
    from tests.scenarios import integration_stub_client
    from tests.simt import emu
    client = integration_stub_client(emu.build(), n_objects=1)
    o = Oracle(1)
    a = np.array([0.3, -0.2, 0.1, -0.4, 0.0, 0.5, 0.2, 0.3, 0.2])
    for _ in range(3):
        client.step(a, True)
        o.step(a)
    assert np.abs(client.joints[0] - o.read()[0]).max() < 1e-9
    r, d, m = o.render()
    assert (client.mask[0] == m).all() and (client.rgb[0] == r).all() and (client.depth[0] == d).all()
    assert np.abs(client.objs[0, 0, :3] - o.get_state()[2][0, :3]).max() < 1e-9


def test_emulated_heavy_solver_kernel_is_the_same_arithmetic():
    """Envs with long Gauss-Seidel schedules (the arm touches something) are solved by k_step_solve_heavy (schedule in shared memory, launched beside
    k_step_solve), the others by k_step_solve in order of decreasing schedule length.
    Same code path for the arithmetic: the state after 25 steps is identical bit for bit with the heavy kernel switched off
    (RR_HEAVY_SOLVER=0) and with a heavy kernel that has room for ONE env (-DHEAVY_GRID=1 -DHEAVY_PER_BLOCK=1: the second heavy env
    of the batch falls to the main kernel, which takes it first); the pressed envs did go through it (schedule length above the heavy
    threshold of 24)."""
    from tests.scenarios import scripted_press
    o, st, cmd = scripted_press(-0.10, 0.50, 0.44)
    rest = Oracle(3).get_state()
    q = np.stack([st[0], rest[0], st[0]]); qd = np.stack([st[1], rest[1], st[1]]); ob = np.stack([st[2], rest[2], st[2]])
    cmds = np.stack([cmd, np.zeros(9), cmd])
    finals, costs = [], []
    for env_off, flags in ((False, ()), (True, ()), (False, ("-DHEAVY_GRID=1", "-DHEAVY_PER_BLOCK=1"))):
        os.environ["RR_HEAVY_SOLVER"] = "0" if env_off else "1"      # 1: always (the default splits a step only when a render followed the last one)
        try:
            sim = EmuSim(3, 3, build_flags=flags)
        finally:
            os.environ.pop("RR_HEAVY_SOLVER", None)
        sim.set_state(q, qd, ob)
        cost = 0
        for t in range(25):
            sim.step(cmds)
            cost = max(cost, int((sim.diag()[0, 3] >> 16) & 0xffff))
        finals.append(sim.get_state()); costs.append(cost)
        sim.close()
    assert costs[0] > 24 and costs[0] == costs[1] == costs[2]
    for other in finals[1:]:
        for a, b in zip(finals[0], other):
            assert np.array_equal(a, b)
    for e in (0, 2):
        assert np.array_equal(finals[0][2][e], finals[0][2][0])        # the same env twice in one batch: same result wherever it ran


def test_emulated_render_right_after_step_takes_the_late_env_path():
    """rr_render called straight after rr_step does not join the heavy-env solver kernel: k_raster walks the envs in the solver's
    order (lightest first), and for the heavy kernel's envs waits for their ticket and computes their matrices itself.  The images
    equal those of a render that joined first (rr_read_state in between) bit for bit, env by env."""
    from tests.scenarios import scripted_press
    o, st, cmd = scripted_press(-0.10, 0.50, 0.44)
    rest = Oracle(3).get_state()
    q = np.stack([rest[0], st[0], rest[0]]); qd = np.stack([rest[1], st[1], rest[1]]); ob = np.stack([rest[2], st[2], rest[2]])
    cmds = np.stack([np.zeros(9), cmd, np.full(9, 0.3)])
    a, b = EmuSim(3, 3), EmuSim(3, 3)
    for sim in (a, b):
        sim.set_state(q, qd, ob)
    for t in range(3):
        a.step(cmds); b.step(cmds)
        ra = a.render()                        # late path (heavy kernel not joined)
        b.read_state(); rb = b.render()        # joined: identity order, k_render_setup's matrices
        assert int((b.diag()[1, 3] >> 16) & 0xffff) > 24          # env 1 is one of the heavy kernel's
        for x, y in zip(ra, rb):
            assert np.array_equal(x, y)
    assert ra[0][1].any() and not np.array_equal(ra[0][0], ra[0][1])


def test_emulated_solver_split_follows_the_call_pattern():
    """rr_step splits the solver over two kernels (and the look-ahead narrowphase over two launches) only when the previous step was
    followed FIRST by a render of all envs — otherwise nothing could hide the heavy kernel.  Counted in kernel launches per rr_step:
    dynamics + solver + look-ahead narrowphase = 3 unsplit, + heavy solver + second narrowphase launch = 5 split."""
    sim = EmuSim(2, 3)
    cmd = np.zeros((2, 9))

    def step_launches():
        n = sim.lib.rr_launch_count(sim.ctx)
        sim.step(cmd)
        return sim.lib.rr_launch_count(sim.ctx) - n
    assert step_launches() == 6          # first step of a context: split by default, and the narrowphase runs in line (nothing looked ahead yet)
    assert step_launches() == 3          # nothing rendered after the first step
    sim.render()
    assert step_launches() == 5          # a render came first
    sim.read_state(); sim.render()
    assert step_launches() == 3          # the state was read before the render: the heavy kernel would have been joined anyway
    sim.render(); sim.read_state()
    assert step_launches() == 5
    sim.render(mask=np.array([1, 0], np.uint8))
    assert step_launches() == 5          # a masked render of all envs counts
    sim.close()
