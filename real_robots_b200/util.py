import numpy as np


def euler_to_quat(rpy):
    """pybullet.getQuaternionFromEuler (btQuaternion::setEulerZYX(yaw, pitch, roll)); returns xyzw."""
    roll, pitch, yaw = rpy
    hy, hp, hr = yaw * 0.5, pitch * 0.5, roll * 0.5
    cy, sy, cp, sp, cr, sr = np.cos(hy), np.sin(hy), np.cos(hp), np.sin(hp), np.cos(hr), np.sin(hr)
    return np.array([sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy, cr * cp * cy + sr * sp * sy])
