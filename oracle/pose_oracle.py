"""CPU ORACLE (test infrastructure, not a product path): fp64 restatement of the
reference's pose algebra, file:line = /root/reference/sim_a_splat/splat/splat_handler.py.

Pinned against the reference's own code by tests/golden/draw_handler.npz, which
tools/make_golden.py produced by executing the unmodified `SplatHandler`
methods (`_load_saved_masks`, `draw_handler`, `get_attached_frame`) in this
container with stub third-party modules.
"""
from __future__ import annotations

import numpy as np


def quat_wxyz_to_R(q) -> np.ndarray:
    q = np.asarray(q, dtype=np.float64)
    w, x, y, z = q / np.sqrt(np.sum(q * q))
    return np.array(
        [
            [w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
            [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
            [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z],
        ]
    )


def R_to_quat_wxyz(R) -> np.ndarray:
    """Eigenvector form (independent of the branchy formula the product uses)."""
    R = np.asarray(R, dtype=np.float64)
    K = np.array(
        [
            [R[0, 0] - R[1, 1] - R[2, 2], 0, 0, 0],
            [R[1, 0] + R[0, 1], R[1, 1] - R[0, 0] - R[2, 2], 0, 0],
            [R[2, 0] + R[0, 2], R[2, 1] + R[1, 2], R[2, 2] - R[0, 0] - R[1, 1], 0],
            [R[1, 2] - R[2, 1], R[2, 0] - R[0, 2], R[0, 1] - R[1, 0], R[0, 0] + R[1, 1] + R[2, 2]],
        ]
    ) / 3.0
    w, v = np.linalg.eigh(K)  # uses the lower triangle
    x, y, z, ww = v[:, np.argmax(w)]
    q = np.array([ww, -x, -y, -z])  # this K yields the conjugate; flip the vector part
    if q[0] < 0:
        q = -q
    return q


def icp_factor(A):
    """:62-83  A = [s*R | t] -> (scale_factor, Ri, ti), with the reference's 1e-6 asserts."""
    A = np.asarray(A, dtype=np.float64)
    cR = A[:3, :3]
    cI = cR.T @ cR
    assert np.all(np.abs(cI[~np.eye(3, dtype=bool)]) < 1e-6)
    s2 = np.mean(np.diag(cI))
    assert np.all(np.abs(np.diag(cI) - s2) < 1e-6)
    s = np.sqrt(s2)
    return float(s), cR / s, A[:3, 3].copy()


def link_transform(q_wxyz, p, weld_t, s, Ri, ti, Rfk, tfk):
    """:250-281  per-link (rot, pos) applied to the link's splat group."""
    q = np.asarray(q_wxyz, dtype=np.float64)
    Rm = quat_wxyz_to_R(q / np.linalg.norm(q))
    tm = np.asarray(p, dtype=np.float64) + np.asarray(weld_t, dtype=np.float64)
    rot = Ri @ Rm @ Rfk.T @ Ri.T
    pos = -rot @ ti - s * (Ri @ Rm @ Rfk.T @ tfk) + s * (Ri @ tm) + ti
    return rot, pos


def mesh_frame(q_wxyz, p, weld_t, s, Ri, ti):
    """:241-263 and :296-310  icp o (q_hat, s*(p + weld_t)) -> (R, t)."""
    q = np.asarray(q_wxyz, dtype=np.float64)
    Rm = quat_wxyz_to_R(q / np.linalg.norm(q))
    t = (np.asarray(p, dtype=np.float64) + np.asarray(weld_t, dtype=np.float64)) * s
    return Ri @ Rm, Ri @ t + ti


def attached_frame(q_wxyz, p, local_t, s, Ri, ti):
    """:316-332  moving-camera pose: icp o (q_hat, s*(p + t_local)); the local
    frame's rotation is ignored and its offset is added in world axes (a
    reference quirk)."""
    return mesh_frame(q_wxyz, p, local_t, s, Ri, ti)


def poses_to_rot12(rots, poss) -> np.ndarray:
    """Pack [..,3,3] and [..,3] into the renderer's fp32 [..,12] (rot row-major | pos)."""
    rots = np.asarray(rots, dtype=np.float64)
    poss = np.asarray(poss, dtype=np.float64)
    return np.concatenate((rots.reshape(*rots.shape[:-2], 9), poss), axis=-1).astype(np.float32)
