"""Independent pins of the parts of the oracle that no reference output can pin (pybullet is absent).

Each test states the arithmetic inside pybullet in a DIFFERENT formulation from the oracle's and requires the two
to agree:

* the retina (TinyRenderer raster + z-buffer, reference env.py:536-567) against a ray caster built from the camera
  parameters alone (eye / target / up / fov of env.py:136,523,548-551) — projection, viewport, integer pixel
  sampling, vertical flip, depth encoding and mask ids;
* the contact solver against conservation laws and Coulomb's law (scene parameters env.py:202-204, friction
  coefficients from the URDFs);
* the free-body integrator against conservation of angular momentum;
* the IK (env.py:372-375) against the forward kinematics.

None of these is a pybullet output: DESIGN.md keeps saying "parity unpinned".
"""
import os

import numpy as np
import pytest

from oracle.orc import Oracle
from real_robots_b200.blob import read_blob

BLOB = os.path.join(os.path.dirname(__file__), "..", "real_robots_b200", "assets", "scene.rrb")
W, H = 320, 240


@pytest.fixture(scope="module")
def blob():
    return read_blob(BLOB)


def _world_triangles(blob, frames):
    tri = blob["tri.data"].astype(np.float64)
    verts, ids = [], []
    for s in range(len(blob["shape.body"])):
        st, ct = int(blob["shape.tri_start"][s]), int(blob["shape.tri_count"][s])
        if ct == 0:
            continue
        R = frames[s, :9].reshape(3, 3)
        verts.append(tri[st:st + ct, :, :3] @ R.T + frames[s, 9:])
        ids.append(np.full(ct, blob["shape.body"][s]))
    return np.concatenate(verts), np.concatenate(ids)


def _ray_cast(V, ID, xs, ys):
    """Nearest triangle along the ray through pixel (x, y) (row 0 = top): Moeller-Trumbore in fp64.

    The camera is rebuilt from the reference's parameters, not from the oracle's matrices.  Pixel x samples
    ndc = 2x/W - 1 (TinyRenderer evaluates barycentrics at integer screen coordinates of a viewport that maps
    ndc [-1, 1] to [0, W]); image rows are TinyRenderer's rows flipped (pybullet copies the image upside down).
    """
    eye = np.array([0.01, 0.0, 1.2]); target = np.array([0.0, 0.0, 0.08]); up = np.array([0.0, 0.0, 1.0])
    f = target - eye; f /= np.linalg.norm(f)
    s = np.cross(f, up); s /= np.linalg.norm(s)
    u = np.cross(s, f)
    th = np.tan(np.radians(80.0) / 2); aspect = W / H
    ndx = 2.0 * xs / W - 1.0
    ndy = 2.0 * (H - 1 - ys) / H - 1.0
    d = f[None] + ndx[:, None] * th * aspect * s[None] + ndy[:, None] * th * u[None]   # depth along f is 1
    v0 = V[:, 0]; e1 = V[:, 1] - v0; e2 = V[:, 2] - v0
    T = eye - v0
    Q = np.cross(T, e1)
    qe2 = (Q * e2).sum(-1)
    t_hit = np.full(len(xs), np.inf); body = np.full(len(xs), -1)
    for i0 in range(0, len(xs), 128):
        dd = d[i0:i0 + 128]
        P = np.cross(dd[:, None, :], e2[None])
        det = (P * e1[None]).sum(-1)
        with np.errstate(divide="ignore", invalid="ignore"):
            inv = 1.0 / np.where(np.abs(det) < 1e-14, np.nan, det)
            uu = (P * T[None]).sum(-1) * inv
            vv = (Q[None] * dd[:, None, :]).sum(-1) * inv
            tt = qe2[None] * inv
            hit = (uu >= 0) & (vv >= 0) & (uu + vv <= 1) & (tt > 0.1)
        tt = np.where(hit, tt, np.inf)
        j = tt.argmin(1)
        tb = tt[np.arange(len(j)), j]
        t_hit[i0:i0 + 128] = tb
        body[i0:i0 + 128] = np.where(np.isfinite(tb), ID[j], -1)
    return t_hit, body


@pytest.mark.parametrize("moved", [False, True])
def test_retina_matches_ray_casting(blob, moved):
    o = Oracle(3)
    if moved:   # arm bent over the table, objects tumbled onto it
        cmd = np.array([0.6, 0.9, -0.4, -1.1, 0.3, 0.8, 0.5, 0.6, 0.9])
        for _ in range(120):
            o.step(cmd)
    rgb, depth, mask = o.render()
    V, ID = _world_triangles(blob, o.shape_frames())
    ys, xs = np.mgrid[2:H:5, (1 if moved else 3):W:5]
    ys = ys.ravel().astype(np.float64); xs = xs.ravel().astype(np.float64)
    t, body = _ray_cast(V, ID, xs, ys)
    m = mask[ys.astype(int), xs.astype(int)]
    # silhouettes: a pixel whose ray grazes an edge may go either way; everything else must agree
    agree = (body == m).mean()
    assert agree >= 0.995, agree
    assert set(np.unique(m).tolist()) >= {-1, 0, 1}
    # depth buffer = OpenGL window depth of the hit (near 0.1, far 100, env.py:548-551), stored in fp32
    near, far = 0.1, 100.0
    ok = (body == m) & (m >= 0)
    d_gl = 0.5 * ((far + near) / (far - near) - 2 * far * near / ((far - near) * t[ok])) + 0.5
    err = np.abs(d_gl - depth[ys.astype(int), xs.astype(int)][ok])
    assert np.median(err) < 1e-7 and np.quantile(err, 0.99) < 1e-5, (np.median(err), err.max())
    # background: the depth formula applied to TinyRenderer's cleared z-buffer, far / (far - near)
    bg = m == -1
    assert (depth[ys.astype(int), xs.astype(int)][bg] == np.float32(far / (far - near))).all()
    assert (rgb[ys.astype(int), xs.astype(int)][bg] == 255).all()


def test_object_collision_conserves_linear_momentum(blob):
    m = blob["obj.mass"]
    o = Oracle(3, gravity=0.0, lin_damping=0.0, ang_damping=0.0)
    _, _, obj = o.get_state()
    obj[:] = 0; obj[:, 6] = 1
    obj[0, :3] = [-0.1, 0.0, 0.7]; obj[0, 7:10] = [0, 0.5, 0]       # cube flies into the resting tomato, mid-air
    obj[1, :3] = [-0.1, 0.2, 0.7]
    obj[2, :3] = [0.3, -0.5, 0.9]
    o.set_state(obj=obj)
    p0 = (m[:, None] * obj[:, 7:10]).sum(0)
    touched = False
    for _ in range(200):
        o.step_sim_only()
        _, _, ob = o.get_state()
        touched |= o.diag()["contacts"] > 0
        assert np.allclose((m[:, None] * ob[:, 7:10]).sum(0), p0, atol=1e-12)
    assert touched
    # restitution 0.1 x 0.01 under Bullet's velocity threshold: perfectly inelastic along the normal
    assert abs(ob[0, 8] - ob[1, 8]) < 1e-9 and abs(ob[0, 8] - 0.75 / 4.5) < 1e-9


def test_sliding_cube_obeys_coulomb_friction():
    # cube (lateral friction 0.5 default) on the table (1.0, table.urdf:5): mu = 0.5 -> deceleration mu g
    o = Oracle(1, lin_damping=0.0, ang_damping=0.0)
    for _ in range(300):
        o.step_sim_only()
    _, _, obj = o.get_state()
    y0 = obj[0, 1]
    obj[0, 7:10] = [0, 1.0, 0]; obj[0, 10:13] = 0
    o.set_state(obj=obj)
    vy = []
    for _ in range(80):
        o.step_sim_only()
        vy.append(o.get_state()[2][0, 8])
    vy = np.array(vy)
    decel = -(vy[30] - vy[5]) / (25 * 0.005)
    assert abs(decel - 0.5 * 9.81) < 0.06 * 0.5 * 9.81, decel
    _, _, ob = o.get_state()
    assert abs((ob[0, 1] - y0) - 1.0 / (2 * 0.5 * 9.81)) < 0.005        # stopping distance v^2 / (2 mu g)
    assert np.abs(ob[0, 7:10]).max() < 1e-3                                # and it has stopped


def test_torque_free_rotation_conserves_angular_momentum(blob):
    I = blob["obj.inertia"][2]                                             # mustard: (0.01, 0.01, 0.001)
    o = Oracle(3, gravity=0.0, lin_damping=0.0, ang_damping=0.0, enable_contacts=0)
    _, _, obj = o.get_state()
    obj[:, 2] += 0.5
    obj[2, 10:13] = [3.0, 0.5, 2.0]
    o.set_state(obj=obj)

    def momentum(b):
        x, y, z, w = b[3:7]
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        return R @ np.diag(I) @ R.T @ b[10:13]

    L0 = momentum(obj[2])
    for _ in range(1000):
        o.step_sim_only()
    ob = o.get_state()[2][2]
    assert abs(np.linalg.norm(ob[3:7]) - 1) < 1e-12
    assert np.linalg.norm(ob[10:13] - obj[2, 10:13]) > 0.3                 # it precesses (gyroscopic term present)
    assert np.linalg.norm(momentum(ob) - L0) < 0.01 * np.linalg.norm(L0)   # explicit integration: 0.3 % in 1000 steps


@pytest.mark.parametrize("pos, quat", [([-0.1, 0.0, 0.6], [0, 1, 0, 0]),
                                       ([-0.2, 0.2, 0.5], [0, 1, 0, 0]),
                                       ([0.0, -0.3, 0.55], [0, 0.7071, 0.7071, 0])])
def test_ik_round_trip_through_forward_kinematics(pos, quat):
    # env.py:372-375: residualThreshold 1e-3, maxNumIterations 1000, end-effector link 7
    o = Oracle(1, enable_contacts=0)
    joints, iters = o.ik(pos, quat)
    assert 0 < iters < 1000
    o.set_state(q=joints, qd=np.zeros(11))
    pose = o.link_pose(7)
    assert np.linalg.norm(pose[:3] - pos) < 1.5e-3
    qn = np.asarray(quat, float) / np.linalg.norm(quat)
    assert abs(abs(pose[3:] @ qn) - 1) < 1e-5


# ---- narrowphase: GJK distance against a quadratic programme, EPA depth against the support-width definition ----

def _hull(blob, s):
    st, ct = int(blob["shape.hull_start"][s]), int(blob["shape.hull_count"][s])
    return blob["hull.verts"][st:st + ct]


def _closest(o, sa, Ra, pa, sb, Rb, pb):
    import ctypes as C
    from oracle.orc import _p
    fa = np.concatenate([Ra.ravel(), pa]); fb = np.concatenate([Rb.ravel(), pb]); out = np.zeros(10)
    assert o.L.orc_test_closest(C.c_void_p(o.h), int(sa), _p(fa), int(sb), _p(fb), _p(out)) == 1
    return out[0], out[1:4], out[4:7], out[7:10]


def _random_pairs(n, seed):
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(seed)
    for _ in range(n):
        sa, sb = rng.choice([19, 20, 21, 9, 10, 12], 2, replace=False)      # cube, tomato, mustard, finger links
        Ra, Rb, Rd = (Rotation.random(random_state=int(rng.integers(1 << 30))).as_matrix() for _ in range(3))
        pa = rng.uniform(-0.1, 0.1, 3)
        yield rng, sa, Ra, pa, sb, Rb, pa + Rd @ np.array([0.25 + 0.1 * rng.random(), 0, 0])


def test_gjk_distance_matches_quadratic_programme(blob):
    """min |sum a_i A_i - sum b_j B_j| over the two vertex simplices, solved by non-negative least squares."""
    from scipy.optimize import nnls
    o = Oracle(3)
    for rng, sa, Ra, pa, sb, Rb, pb in _random_pairs(16, 0):
        d, n, wa, wb = _closest(o, sa, Ra, pa, sb, Rb, pb)
        # also a near-touching configuration: slide B along the separating normal until 1 mm remains
        for gap in (None, 1e-3):
            pb_k = pb if gap is None else pb + n * (d - gap) * (1 if (pb - pa) @ n < 0 else -1)
            dk, nk, wa, wb = _closest(o, sa, Ra, pa, sb, Rb, pb_k)
            A = _hull(blob, sa) @ Ra.T + pa; B = _hull(blob, sb) @ Rb.T + pb_k
            w = 1e3
            M = np.zeros((5, len(A) + len(B)))
            M[:3, :len(A)] = A.T; M[:3, len(A):] = -B.T; M[3, :len(A)] = w; M[4, len(A):] = w
            x, _ = nnls(M, np.array([0, 0, 0, w, w]), maxiter=20000)
            a = x[:len(A)] / x[:len(A)].sum(); b = x[len(A):] / x[len(A):].sum()
            d_qp = np.linalg.norm(A.T @ a - B.T @ b)
            assert abs(dk - d_qp) < 1e-9, (sa, sb, dk, d_qp)
            assert abs(np.linalg.norm(wa - wb) - dk) < 1e-12 and abs(np.linalg.norm(nk) - 1) < 1e-12
            if gap is not None:
                assert abs(dk - gap) < 1e-9


def test_epa_depth_is_the_minimum_support_width(blob):
    """Penetration depth = min over unit n of max_{a in A, b in B} n.(a - b); EPA must attain it along its normal and
    no hull-face normal or sampled direction may do better."""
    o = Oracle(3)
    for rng, sa, Ra, pa, sb, Rb, pb in _random_pairs(12, 1):
        d0, n0, _, _ = _closest(o, sa, Ra, pa, sb, Rb, pb)
        push = 0.004
        pb2 = pb + n0 * (d0 + push) * (1 if (pb - pa) @ n0 < 0 else -1)       # overlap by at most `push`
        d, n, wa, wb = _closest(o, sa, Ra, pa, sb, Rb, pb2)
        assert -push - 1e-12 <= d < 0
        A = _hull(blob, sa) @ Ra.T + pa; B = _hull(blob, sb) @ Rb.T + pb2
        width = lambda c: (A @ c.T).max(0) - (B @ c.T).min(0)
        assert min(abs(width(n[None])[0] + d), abs(width(-n[None])[0] + d)) < 1e-12
        fa = blob["hull.faces"][int(blob["shape.face_start"][sa]):][:int(blob["shape.face_count"][sa]), :3] @ Ra.T
        fb = blob["hull.faces"][int(blob["shape.face_start"][sb]):][:int(blob["shape.face_count"][sb]), :3] @ Rb.T
        cands = np.concatenate([fa, -fb, rng.normal(size=(20000, 3))])
        cands /= np.linalg.norm(cands, axis=1, keepdims=True)
        assert width(cands).min() >= -d - 1e-9


def test_resting_contact_impulses_carry_the_weight(blob):
    """Statics: the normal impulses of each object's contact points sum to m g dt per step, along +z — to 1e-9 without the
    torsional rows.  With them (the default) the tomato, which rests on two points of a round hull, keeps creeping at
    ~5e-4 rad/s (the rolling rows hold against the penetration correction of the unequal contact depths; PGS settles on a
    blend instead of converging), so there the balance is held to 1e-4 of the weight."""
    for tf, tol, vmax in ((0, 1e-9, 1e-4), (1, 1e-4 * 9.81 * 0.005, 1e-3)):
        o = Oracle(3, torsional_friction=tf)
        for _ in range(400):
            o.step_sim_only()
        a, b, d = o.contacts()                                  # [n][11]: worldA3 worldB3 normal3 dist impulse
        for shape, m in zip(blob["obj.shape"], blob["obj.mass"]):
            sel = (a == shape) | (b == shape)
            assert 2 <= sel.sum() <= 4
            assert set(b[sel].tolist()) <= set(blob["table.shape"].tolist())
            assert np.allclose(d[sel, 6:9], [0, 0, 1], atol=1e-9)
            assert abs(d[sel, 10].sum() - m * 9.81 * 0.005) < tol * max(1.0, m) and (d[sel, 10] >= 0).all()
        assert np.abs(o.get_state()[2][:, 7:13]).max() < vmax


def test_spinning_friction_decelerates_at_the_torsional_rate(blob):
    """Torsional friction rows (urdf/cube.urdf:6-7: spinning_friction 1e-4; Bullet adds one spinning row per manifold point when
    the combined coefficient spin_obj * friction_table is > 0, limits +-(coefficient x normal impulse)).  A cube spinning flat
    on the table is braked by lateral friction at its four contact points AND by the spinning rows; Coulomb friction does
    not depend on the speed, so the difference between a run with and without the torsional rows is the torsional torque
    alone: coefficient * m * g about the vertical axis, i.e. an extra angular deceleration of coefficient * m * g / I_zz."""
    runs = {}
    for tf in (1, 0):
        o = Oracle(3, torsional_friction=tf)
        for _ in range(150):
            o.step(np.zeros(9))
        q, qd, ob = o.get_state()
        ob[0, 7:] = 0
        ob[0, 12] = 25.0                                  # rad/s about z
        o.set_state(q, qd, ob)
        w = []
        for _ in range(10):
            o.step(np.zeros(9))
            w.append(o.get_state()[2][0, 12])
        runs[tf] = np.array(w)
    assert runs[1][-1] > 10 and runs[0][-1] > 10          # both still spinning: friction is saturated throughout
    extra = np.diff(runs[0] - runs[1])[1:8].mean()        # extra loss of angular velocity per step with the torsional rows
    coef = blob["shape.spinning"][19] * blob["shape.friction"][17]      # cube x table_base
    pred = coef * blob["obj.mass"][0] * 9.81 / blob["obj.inertia"][0][2] * 0.005
    assert abs(extra - pred) < 0.1 * pred, (extra, pred)
