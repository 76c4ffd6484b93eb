"""Triangle meshes -> Gaussians.

The reference shows the PushT T-block and the robot links as viser TRIANGLE MESHES drawn over the splat
(/root/reference/sim_a_splat/splat/splat_handler.py:190-194, :199-219, moved at :241-249, :296-310); the
browser rasterises them.  This renderer draws Gaussians only, so a mesh that must be visible in the
observation (the T-block is the whole point of a PushT image) is turned into surface Gaussians once, at
load time: flat discs sampled uniformly over the triangles, oriented by the face normal, coloured with the
mesh colour (SH degree 0).  (The reference's own `create_gs_mesh`, ellipsoids/mesh_utils.py:4-30, goes the
other way -- Gaussians to an ellipsoid mesh for export -- and is out of scope.)
"""
from __future__ import annotations

from typing import Tuple

import numpy as np

from ..scene import C0


class TriMesh:
    """vertices [V,3] float64, faces [F,3] int -- the two attributes of trimesh.Trimesh this module needs."""

    def __init__(self, vertices, faces):
        self.vertices = np.asarray(vertices, dtype=np.float64).reshape(-1, 3)
        self.faces = np.asarray(faces, dtype=np.int64).reshape(-1, 3)


def load_obj(path: str) -> TriMesh:
    """Minimal Wavefront OBJ reader: `v x y z` and `f a/b/c ...` (polygons are fan-triangulated)."""
    verts, faces = [], []
    with open(path, "r") as f:
        for line in f:
            p = line.split()
            if not p:
                continue
            if p[0] == "v":
                verts.append([float(p[1]), float(p[2]), float(p[3])])
            elif p[0] == "f":
                idx = [int(tok.split("/")[0]) for tok in p[1:]]
                idx = [i - 1 if i > 0 else len(verts) + i for i in idx]
                for k in range(1, len(idx) - 1):
                    faces.append([idx[0], idx[k], idx[k + 1]])
    if not verts or not faces:
        raise ValueError(f"{path}: no triangles found")
    return TriMesh(verts, faces)


def _quat_from_z_to(n: np.ndarray) -> np.ndarray:
    """wxyz quaternions rotating +z onto the unit vectors n [M,3]."""
    z = np.array([0.0, 0.0, 1.0])
    c = n @ z
    axis = np.cross(np.broadcast_to(z, n.shape), n)
    q = np.concatenate([(1.0 + c)[:, None], axis], axis=1)
    flip = c < -1.0 + 1e-9          # n = -z: rotate pi about x
    q[flip] = np.array([0.0, 1.0, 0.0, 0.0])
    return q / np.linalg.norm(q, axis=1, keepdims=True)


def mesh_to_gaussians(mesh, n: int, rgb=(0.956, 0.396, 0.365), seed: int = 0, opacity: float = 0.97,
                      thickness: float = 0.15) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """n surface Gaussians: (means [n,3], quats wxyz [n,4], scales [n,3], opacities [n], sh [n,1,3]).
    In-plane sigma = half the mean sample spacing x 1.5 (neighbouring discs overlap into an opaque sheet);
    sigma along the normal = thickness x that."""
    v, f = np.asarray(mesh.vertices, dtype=np.float64), np.asarray(mesh.faces, dtype=np.int64)
    a, b, c = v[f[:, 0]], v[f[:, 1]], v[f[:, 2]]
    cr = np.cross(b - a, c - a)
    area = 0.5 * np.linalg.norm(cr, axis=1)
    keep = area > 0
    a, b, c, cr, area = a[keep], b[keep], c[keep], cr[keep], area[keep]
    rng = np.random.default_rng(seed)
    tri = rng.choice(area.shape[0], size=n, p=area / area.sum())
    r1, r2 = np.sqrt(rng.uniform(size=n)), rng.uniform(size=n)
    pts = (1 - r1)[:, None] * a[tri] + (r1 * (1 - r2))[:, None] * b[tri] + (r1 * r2)[:, None] * c[tri]
    normals = cr[tri] / (2.0 * area[tri])[:, None]
    sigma = 0.75 * np.sqrt(area.sum() / n)
    scales = np.tile(np.array([sigma, sigma, thickness * sigma]), (n, 1))
    sh = np.tile(((np.asarray(rgb, dtype=np.float64) - 0.5) / C0)[None, None, :], (n, 1, 1))
    return (pts.astype(np.float32), _quat_from_z_to(normals).astype(np.float32), scales.astype(np.float32),
            np.full(n, opacity, dtype=np.float32), sh.astype(np.float32))
