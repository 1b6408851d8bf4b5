"""pybullet camera helpers in float32 (computeViewMatrix, computeViewMatrixFromYawPitchRoll, computeProjectionMatrixFOV)
as the reference uses them (env.py:489-500, 543-551).  Column-major float32[16], like pybullet returns them."""
import numpy as np

f32 = np.float32


def _norm(v):
    return (v / f32(np.sqrt(f32(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])))).astype(f32)


def view_from_positions(eye, target, up):
    eye, target, up = (np.asarray(a, f32) for a in (eye, target, up))
    f = _norm(target - eye)
    u = _norm(up)
    s = _norm(np.cross(f, u).astype(f32))
    u = np.cross(s, f).astype(f32)
    m = np.zeros(16, f32)
    m[0], m[4], m[8] = s
    m[1], m[5], m[9] = u
    m[2], m[6], m[10] = -f
    m[12] = -f32(np.dot(s, eye)); m[13] = -f32(np.dot(u, eye)); m[14] = f32(np.dot(f, eye)); m[15] = 1
    return m


def view_from_yaw_pitch_roll(target, distance, yaw, pitch, roll, up_axis=2):
    """b3ComputeViewMatrixFromYawPitchRoll (roll is ignored by pybullet)."""
    yaw_r, pitch_r = f32(yaw) * f32(0.01745329251994329547), f32(pitch) * f32(0.01745329251994329547)
    from .util import euler_to_quat
    if up_axis == 2:
        forward_axis, cam_up = 1, np.array([0, 0, 1], f32)
        q = euler_to_quat([pitch_r, 0.0, yaw_r])        # setEulerZYX(yaw, roll=0, pitch): z = yaw, y = 0, x = pitch
    else:
        forward_axis, cam_up = 2, np.array([0, 1, 0], f32)
        q = euler_to_quat([-pitch_r, yaw_r, 0.0])
    x, y, z, w = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                  [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                  [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]], f32)
    eye = np.zeros(3, f32)
    eye[forward_axis] = -f32(distance)
    eye = (R @ eye).astype(f32)
    up = (R @ cam_up).astype(f32)
    tgt = np.asarray(target, f32)
    return view_from_positions(eye + tgt, tgt, up)


def projection_fov(fov, aspect, near, far):
    y = f32(1.0 / np.tan((np.pi / 180.0) * float(f32(fov)) / 2))
    x = f32(y / f32(aspect))
    near, far = f32(near), f32(far)
    m = np.zeros(16, f32)
    m[0], m[5] = x, y
    m[10] = (near + far) / (near - far)
    m[11] = -1
    m[14] = (f32(2) * far * near) / (near - far)
    return m
