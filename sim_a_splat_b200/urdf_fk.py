"""Minimal URDF forward kinematics of the *visual meshes* -- what the reference gets from
`urchin.URDF.load(...).visual_trimesh_fk(cfg)` (/root/reference/sim_a_splat/splat/splat_handler.py:156-163,
:197): one world transform per visual mesh, in link order then visual order.  urchin / trimesh are not
available offline; only the transforms are needed on the render path (`fk_tf`), not the mesh geometry.
"""
from __future__ import annotations

import xml.etree.ElementTree as ET
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np


def _rpy_xyz(el) -> np.ndarray:
    T = np.eye(4)
    if el is None:
        return T
    xyz = [float(v) for v in el.get("xyz", "0 0 0").split()]
    r, p, y = (float(v) for v in el.get("rpy", "0 0 0").split())
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    T[:3, :3] = np.array([[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
                          [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
                          [-sp, cp * sr, cp * cr]])
    T[:3, 3] = xyz
    return T


def _axis_angle(axis: np.ndarray, q: float) -> np.ndarray:
    a = axis / (np.linalg.norm(axis) or 1.0)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    T = np.eye(4)
    T[:3, :3] = np.eye(3) + np.sin(q) * K + (1 - np.cos(q)) * (K @ K)
    return T


@dataclass
class Visual:
    link: str
    origin: np.ndarray
    mesh_filename: Optional[str]
    rgba: Optional[np.ndarray]


@dataclass
class Joint:
    name: str
    jtype: str
    parent: str
    child: str
    origin: np.ndarray
    axis: np.ndarray
    mimic: Optional[tuple] = None


@dataclass
class SimpleURDF:
    links: List[str] = field(default_factory=list)
    visuals: List[Visual] = field(default_factory=list)
    joints: List[Joint] = field(default_factory=list)

    @staticmethod
    def load(path: str) -> "SimpleURDF":
        root = ET.parse(path).getroot()
        materials = {}
        for m in root.findall("material"):
            c = m.find("color")
            if c is not None:
                materials[m.get("name")] = np.array([float(v) for v in c.get("rgba").split()])
        u = SimpleURDF()
        for l in root.findall("link"):
            u.links.append(l.get("name"))
            for v in l.findall("visual"):
                geom = v.find("geometry")
                mesh = geom.find("mesh") if geom is not None else None
                mat = v.find("material")
                rgba = None
                if mat is not None:
                    c = mat.find("color")
                    rgba = np.array([float(x) for x in c.get("rgba").split()]) if c is not None else materials.get(mat.get("name"))
                u.visuals.append(Visual(l.get("name"), _rpy_xyz(v.find("origin")), mesh.get("filename") if mesh is not None else None, rgba))
        for j in root.findall("joint"):
            ax = j.find("axis")
            mim = j.find("mimic")
            u.joints.append(Joint(j.get("name"), j.get("type"), j.find("parent").get("link"), j.find("child").get("link"),
                                  _rpy_xyz(j.find("origin")),
                                  np.array([float(v) for v in ax.get("xyz").split()]) if ax is not None else np.array([1.0, 0, 0]),
                                  (mim.get("joint"), float(mim.get("multiplier", 1)), float(mim.get("offset", 0))) if mim is not None else None))
        return u

    @property
    def actuated_joint_names(self) -> List[str]:
        return [j.name for j in self.joints if j.jtype in ("revolute", "continuous", "prismatic") and j.mimic is None]

    def link_fk(self, cfg: Optional[Dict[str, float]] = None) -> Dict[str, np.ndarray]:
        cfg = dict(cfg or {})
        children = {j.child for j in self.joints}
        T = {l: np.eye(4) for l in self.links if l not in children}  # roots
        pending = list(self.joints)
        while pending:
            progressed = False
            for j in list(pending):
                if j.parent in T:
                    q = cfg.get(j.name, 0.0)
                    if j.mimic is not None:
                        q = cfg.get(j.mimic[0], 0.0) * j.mimic[1] + j.mimic[2]
                    M = np.eye(4)
                    if j.jtype in ("revolute", "continuous"):
                        M = _axis_angle(j.axis, q)
                    elif j.jtype == "prismatic":
                        M[:3, 3] = j.axis / (np.linalg.norm(j.axis) or 1.0) * q
                    T[j.child] = T[j.parent] @ j.origin @ M
                    pending.remove(j)
                    progressed = True
            if not progressed:
                raise ValueError("URDF joint graph is not a tree rooted at a link without parent")
        return T

    def visual_mesh_fk(self, cfg: Optional[Dict[str, float]] = None) -> List[np.ndarray]:
        """World transform of every visual *mesh*, link order then visual order (== list(fk.values()) of
        urchin's visual_trimesh_fk for single-geometry meshes)."""
        T = self.link_fk(cfg)
        return [T[v.link] @ v.origin for v in self.visuals if v.mesh_filename is not None]

    def mesh_visuals(self) -> List[Visual]:
        return [v for v in self.visuals if v.mesh_filename is not None]


def _axis_angle_batch(axis: np.ndarray, q: np.ndarray) -> np.ndarray:
    a = axis / (np.linalg.norm(axis) or 1.0)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    T = np.tile(np.eye(4), (q.shape[0], 1, 1))
    T[:, :3, :3] = np.eye(3) + np.sin(q)[:, None, None] * K + (1 - np.cos(q))[:, None, None] * (K @ K)
    return T


class BatchedBodyPoseSource:
    """Body poses for E envs at once from joint configurations -- the pose source of SURVEY.md 8f-1 for users
    without Drake: the draw message of the reference (`ManipulatorSimEnv._generate_draw_msg`,
    /root/reference/sim_a_splat/env/manipulator/manipulator_env.py:377-390) carries the world pose of every
    robot body; here the same [E, B, 7] block (quaternion wxyz | position) comes from the URDF's forward
    kinematics, vectorised over envs, for the bodies that carry visual meshes in URDF order -- the order in
    which `SplatHandler` pairs bodies with link groups.  Feed it to `SplatRenderer.set_link_poses_from_bodies`
    (with `SplatHandler.pose_model()`); a free body (the T-block) is appended by the caller."""

    def __init__(self, urdf: "SimpleURDF", base: Optional[np.ndarray] = None):
        self.urdf = urdf
        self.base = np.eye(4) if base is None else np.asarray(base, dtype=np.float64)
        self.joint_names = urdf.actuated_joint_names
        self.bodies = [v.link for v in urdf.mesh_visuals()]
        children = {j.child for j in urdf.joints}
        self._roots = [l for l in urdf.links if l not in children]
        order, known, pending = [], set(self._roots), list(urdf.joints)
        while pending:  # parent-before-child joint order, resolved once
            nxt = [j for j in pending if j.parent in known]
            if not nxt:
                raise ValueError("URDF joint graph is not a tree rooted at a link without parent")
            for j in nxt:
                order.append(j)
                known.add(j.child)
                pending.remove(j)
        self._order = order

    def link_fk(self, q: np.ndarray) -> Dict[str, np.ndarray]:
        """q [E, n_actuated] -> {link: [E,4,4]} world transforms."""
        q = np.asarray(q, dtype=np.float64).reshape(-1, len(self.joint_names))
        E = q.shape[0]
        col = {n: k for k, n in enumerate(self.joint_names)}
        T = {l: np.tile(self.base, (E, 1, 1)) for l in self._roots}
        for j in self._order:
            if j.mimic is not None:
                qj = q[:, col[j.mimic[0]]] * j.mimic[1] + j.mimic[2] if j.mimic[0] in col else np.zeros(E)
            else:
                qj = q[:, col[j.name]] if j.name in col else np.zeros(E)
            if j.jtype in ("revolute", "continuous"):
                M = _axis_angle_batch(j.axis, qj)
            else:
                M = np.tile(np.eye(4), (E, 1, 1))
                if j.jtype == "prismatic":
                    M[:, :3, 3] = (j.axis / (np.linalg.norm(j.axis) or 1.0))[None, :] * qj[:, None]
            T[j.child] = T[j.parent] @ j.origin @ M
        return T

    def device_model(self, n_links: Optional[int] = None) -> Dict[str, np.ndarray]:
        """The joint tree as flat arrays for `SplatRenderer.set_fk_model` (sas_set_fk_model): joints in
        parent-before-child order; link frame 0 = the root link(s) at `base`, frame j+1 = child link of joint j."""
        frame = {l: 0 for l in self._roots}
        for j, jt in enumerate(self._order):
            frame[jt.child] = j + 1
        col = {n: k for k, n in enumerate(self.joint_names)}
        J = len(self._order)
        parent = np.array([frame[j.parent] for j in self._order], dtype=np.int32)
        jtype = np.array([1 if j.jtype in ("revolute", "continuous") else 2 if j.jtype == "prismatic" else 0 for j in self._order], dtype=np.int32)
        qindex = np.full(J, -1, dtype=np.int32)
        mimic = np.tile(np.array([1.0, 0.0]), (J, 1))
        for k, j in enumerate(self._order):
            if jtype[k] == 0:
                continue
            if j.mimic is not None:
                qindex[k] = col.get(j.mimic[0], -1)
                mimic[k] = (j.mimic[1], j.mimic[2])
                if qindex[k] < 0:      # follows a joint that is not actuated: the host path uses q = 0 (not m*0 + o)
                    mimic[k] = (1.0, 0.0)
            else:
                qindex[k] = col.get(j.name, -1)
        axis = np.stack([j.axis / (np.linalg.norm(j.axis) or 1.0) for j in self._order]).astype(np.float64)
        bodies = self.bodies if n_links is None else self.bodies[:n_links]
        return dict(parent=parent, type=jtype, qindex=qindex, mimic=np.ascontiguousarray(mimic, dtype=np.float64),
                    origin=np.ascontiguousarray(np.stack([j.origin for j in self._order]), dtype=np.float64), axis=np.ascontiguousarray(axis),
                    base=np.ascontiguousarray(self.base, dtype=np.float64), body_frame=np.array([frame[b] for b in bodies], dtype=np.int32),
                    n_actuated=len(self.joint_names))

    def body_poses(self, q: np.ndarray, n_links: Optional[int] = None) -> np.ndarray:
        """[E, B, 7] fp64 (wxyz | xyz) of the mesh-carrying bodies (first n_links of them)."""
        from .transforms import SO3

        T = self.link_fk(q)
        bodies = self.bodies if n_links is None else self.bodies[:n_links]
        E = next(iter(T.values())).shape[0]
        out = np.empty((E, len(bodies), 7), dtype=np.float64)
        for b, name in enumerate(bodies):
            out[:, b, 4:] = T[name][:, :3, 3]
            out[:, b, :4] = np.stack([SO3.from_matrix(R).wxyz for R in T[name][:, :3, :3]])
        return out


def visual_fk_from_config(urdf_path: str, joint_config: Sequence[float]) -> List[np.ndarray]:
    u = SimpleURDF.load(urdf_path)
    return u.visual_mesh_fk(dict(zip(u.actuated_joint_names, [float(v) for v in joint_config])))
