"""A stand-in for `viser.transforms` (SO3 / SE3, the part /root/reference/sim_a_splat/splat/splat_handler.py touches)
built ONLY on scipy.spatial.transform.Rotation -- deliberately independent of sim_a_splat_b200/transforms.py, so that the
golden vectors tools/make_golden.py produces by executing the reference's draw_handler do not run through this
repo's own quaternion code.  Used by tools/make_golden.py only."""
from __future__ import annotations

import numpy as np
from scipy.spatial.transform import Rotation as _R


def _rot_from_wxyz(wxyz) -> _R:
    w, x, y, z = (float(v) for v in wxyz)
    return _R.from_quat([x, y, z, w])


class SO3:
    def __init__(self, wxyz):
        self.wxyz = np.asarray(wxyz, dtype=np.float64)

    @staticmethod
    def from_matrix(m) -> "SO3":
        x, y, z, w = _R.from_matrix(np.asarray(m, dtype=np.float64)).as_quat()
        return SO3(np.array([w, x, y, z]))

    @staticmethod
    def from_x_radians(a: float) -> "SO3":
        x, y, z, w = _R.from_euler("x", float(a)).as_quat()
        return SO3(np.array([w, x, y, z]))

    @staticmethod
    def exp(v) -> "SO3":
        x, y, z, w = _R.from_rotvec(np.asarray(v, dtype=np.float64)).as_quat()
        return SO3(np.array([w, x, y, z]))

    def as_matrix(self) -> np.ndarray:
        return _rot_from_wxyz(self.wxyz).as_matrix()


class SE3:
    def __init__(self, wxyz_xyz):
        self.wxyz_xyz = np.asarray(wxyz_xyz, dtype=np.float64)

    @staticmethod
    def from_rotation_and_translation(rotation: SO3, translation) -> "SE3":
        return SE3(np.concatenate((rotation.wxyz, np.asarray(translation, dtype=np.float64))))

    @staticmethod
    def from_matrix(m) -> "SE3":
        m = np.asarray(m, dtype=np.float64)
        return SE3.from_rotation_and_translation(SO3.from_matrix(m[:3, :3]), m[:3, 3])

    def rotation(self) -> SO3:
        return SO3(self.wxyz_xyz[:4])

    def translation(self) -> np.ndarray:
        return self.wxyz_xyz[4:].copy()

    def multiply(self, other: "SE3") -> "SE3":
        Ra, Rb = _rot_from_wxyz(self.wxyz_xyz[:4]), _rot_from_wxyz(other.wxyz_xyz[:4])
        x, y, z, w = (Ra * Rb).as_quat()
        return SE3(np.concatenate(([w, x, y, z], Ra.apply(other.wxyz_xyz[4:]) + self.wxyz_xyz[4:])))

    __matmul__ = multiply
