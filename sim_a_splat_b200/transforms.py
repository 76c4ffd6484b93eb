"""Minimal SO(3)/SE(3) types with the `viser.transforms` surface the reference uses.

The reference imports `viser.transforms as tf` and touches exactly this subset
(/root/reference/sim_a_splat/splat/splat_handler.py:77-83, :241-281, :316-332;
/root/reference/sim_a_splat/env/splat/splat_env_wrapper.py:57-64, :117, :155-156;
/root/reference/tests/test_splat_wrapper.py:77-101):

    tf.SO3(wxyz), tf.SO3.from_matrix, tf.SO3.from_x_radians, .wxyz, .as_matrix(),
    tf.SE3(wxyz_xyz=...), tf.SE3.from_rotation_and_translation, tf.SE3.from_matrix,
    .rotation(), .translation(), .multiply(), .inverse(), .as_matrix(), @

Quaternions are wxyz everywhere; everything is float64 numpy on the host.
"""
from __future__ import annotations

import numpy as np


def _quat_to_matrix(wxyz: np.ndarray) -> np.ndarray:
    w, x, y, z = (float(v) for v in wxyz)
    n = w * w + x * x + y * y + z * z
    s = 2.0 / n
    return np.array(
        [
            [1.0 - s * (y * y + z * z), s * (x * y - z * w), s * (x * z + y * w)],
            [s * (x * y + z * w), 1.0 - s * (x * x + z * z), s * (y * z - x * w)],
            [s * (x * z - y * w), s * (y * z + x * w), 1.0 - s * (x * x + y * y)],
        ],
        dtype=np.float64,
    )


def _matrix_to_quat(m: np.ndarray) -> np.ndarray:
    """Branch on the largest of (trace, m00, m11, m22) so the divisor stays away from 0."""
    m = np.asarray(m, dtype=np.float64)
    t = m[0, 0] + m[1, 1] + m[2, 2]
    if t > 0.0:
        r = np.sqrt(1.0 + t)
        q = np.array([0.5 * r, (m[2, 1] - m[1, 2]) / (2 * r), (m[0, 2] - m[2, 0]) / (2 * r), (m[1, 0] - m[0, 1]) / (2 * r)])
    elif m[0, 0] >= m[1, 1] and m[0, 0] >= m[2, 2]:
        r = np.sqrt(1.0 + m[0, 0] - m[1, 1] - m[2, 2])
        q = np.array([(m[2, 1] - m[1, 2]) / (2 * r), 0.5 * r, (m[0, 1] + m[1, 0]) / (2 * r), (m[0, 2] + m[2, 0]) / (2 * r)])
    elif m[1, 1] >= m[2, 2]:
        r = np.sqrt(1.0 - m[0, 0] + m[1, 1] - m[2, 2])
        q = np.array([(m[0, 2] - m[2, 0]) / (2 * r), (m[0, 1] + m[1, 0]) / (2 * r), 0.5 * r, (m[1, 2] + m[2, 1]) / (2 * r)])
    else:
        r = np.sqrt(1.0 - m[0, 0] - m[1, 1] + m[2, 2])
        q = np.array([(m[1, 0] - m[0, 1]) / (2 * r), (m[0, 2] + m[2, 0]) / (2 * r), (m[1, 2] + m[2, 1]) / (2 * r), 0.5 * r])
    return q / np.linalg.norm(q)


class SO3:
    def __init__(self, wxyz):
        self.wxyz = np.asarray(wxyz, dtype=np.float64).reshape(4)

    @staticmethod
    def identity() -> "SO3":
        return SO3(np.array([1.0, 0.0, 0.0, 0.0]))

    @staticmethod
    def from_matrix(matrix) -> "SO3":
        return SO3(_matrix_to_quat(np.asarray(matrix, dtype=np.float64)))

    @staticmethod
    def _from_axis_radians(axis: int, theta: float) -> "SO3":
        q = np.zeros(4)
        q[0] = np.cos(theta / 2.0)
        q[1 + axis] = np.sin(theta / 2.0)
        return SO3(q)

    @staticmethod
    def from_x_radians(theta: float) -> "SO3":
        return SO3._from_axis_radians(0, theta)

    @staticmethod
    def from_y_radians(theta: float) -> "SO3":
        return SO3._from_axis_radians(1, theta)

    @staticmethod
    def from_z_radians(theta: float) -> "SO3":
        return SO3._from_axis_radians(2, theta)

    @staticmethod
    def exp(omega) -> "SO3":
        omega = np.asarray(omega, dtype=np.float64).reshape(3)
        th = float(np.linalg.norm(omega))
        if th < 1e-12:
            q = np.concatenate(([1.0], 0.5 * omega))
            return SO3(q / np.linalg.norm(q))
        return SO3(np.concatenate(([np.cos(th / 2)], np.sin(th / 2) * omega / th)))

    def as_matrix(self) -> np.ndarray:
        return _quat_to_matrix(self.wxyz)

    def inverse(self) -> "SO3":
        w, x, y, z = self.wxyz
        return SO3(np.array([w, -x, -y, -z]) / float(self.wxyz @ self.wxyz))

    def multiply(self, other: "SO3") -> "SO3":
        w0, x0, y0, z0 = self.wxyz
        w1, x1, y1, z1 = other.wxyz
        return SO3(
            np.array(
                [
                    w0 * w1 - x0 * x1 - y0 * y1 - z0 * z1,
                    w0 * x1 + x0 * w1 + y0 * z1 - z0 * y1,
                    w0 * y1 - x0 * z1 + y0 * w1 + z0 * x1,
                    w0 * z1 + x0 * y1 - y0 * x1 + z0 * w1,
                ]
            )
        )

    def apply(self, v) -> np.ndarray:
        return self.as_matrix() @ np.asarray(v, dtype=np.float64)

    def __matmul__(self, other):
        if isinstance(other, SO3):
            return self.multiply(other)
        return self.apply(other)

    def __repr__(self) -> str:
        return f"SO3(wxyz={np.round(self.wxyz, 5)})"


class SE3:
    def __init__(self, wxyz_xyz):
        self.wxyz_xyz = np.asarray(wxyz_xyz, dtype=np.float64).reshape(7)

    @staticmethod
    def identity() -> "SE3":
        return SE3(np.array([1.0, 0, 0, 0, 0, 0, 0]))

    @staticmethod
    def from_rotation_and_translation(rotation: SO3, translation) -> "SE3":
        return SE3(np.concatenate((rotation.wxyz, np.asarray(translation, dtype=np.float64).reshape(3))))

    @staticmethod
    def from_matrix(matrix) -> "SE3":
        matrix = np.asarray(matrix, dtype=np.float64)
        return SE3.from_rotation_and_translation(SO3.from_matrix(matrix[:3, :3]), matrix[:3, 3])

    def rotation(self) -> SO3:
        return SO3(self.wxyz_xyz[:4])

    def translation(self) -> np.ndarray:
        return self.wxyz_xyz[4:].copy()

    def as_matrix(self) -> np.ndarray:
        m = np.eye(4)
        m[:3, :3] = self.rotation().as_matrix()
        m[:3, 3] = self.wxyz_xyz[4:]
        return m

    def multiply(self, other: "SE3") -> "SE3":
        r = self.rotation()
        return SE3.from_rotation_and_translation(r.multiply(other.rotation()), r.apply(other.translation()) + self.translation())

    def inverse(self) -> "SE3":
        ri = self.rotation().inverse()
        return SE3.from_rotation_and_translation(ri, -ri.apply(self.translation()))

    def apply(self, v) -> np.ndarray:
        return self.rotation().apply(v) + self.translation()

    def __matmul__(self, other):
        if isinstance(other, SE3):
            return self.multiply(other)
        return self.apply(other)

    def __repr__(self) -> str:
        return f"SE3(wxyz={np.round(self.wxyz_xyz[:4], 5)}, xyz={np.round(self.wxyz_xyz[4:], 5)})"
