#!/usr/bin/env python
"""Generate tests/golden/*.npz from the reference itself (run in the build container only).

`/root/reference` is Python over third-party packages that are not installed
(viser, pydrake, urchin, open3d, trimesh, nerfstudio, gsplat).  The pieces of
the hot path that ARE the reference's own arithmetic are executed here,
unmodified, by importing the reference's files with stub modules in
`sys.modules` for the absent third-party imports:

* compute_cov.npz   - sim_a_splat/ellipsoids/covariance_utils.py (pure torch; runs as is)
* sh2rgb.npz        - `C0` and `SH2RGB` lifted by `ast` from ns_utils/nerfstudio_utils.py:43-47
                      (the module itself imports nerfstudio at the top)
* draw_handler.npz  - `SplatHandler._load_saved_masks`, `.draw_handler`,
                      `.get_attached_frame` from splat/splat_handler.py, bound to a bare
                      object; `viser.transforms` is replaced by a
                      stand-in built only on scipy's Rotation (tools/viser_tf_scipy.py), NOT by
                      this repo's own transforms.py
* assets.npz        - the real masks (bit-packed), ICP matrices, intrinsics, a few
                      training camera poses, dataparser transforms and a subsample of
                      the robot surface points, so the GPU box (which has no
                      /root/reference) can build scenes with the real N / mask structure.

Nothing under tests/, bench.py or smoke() reads /root/reference at run time.
"""
from __future__ import annotations

import ast
import importlib.util
import json
import sys
import types
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parents[1]
REF = Path("/root/reference")
OUT = REPO / "tests" / "golden"
sys.path.insert(0, str(REPO))


def _load_by_path(name: str, path: Path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def make_compute_cov():
    cu = _load_by_path("ref_covariance_utils", REF / "sim_a_splat/ellipsoids/covariance_utils.py")
    g = torch.Generator().manual_seed(7)
    quats = torch.randn(512, 4, generator=g)
    quats[:64] *= 3.7  # un-normalised, as stored by splatfacto
    quats[64:72] = torch.tensor([1.0, 0.0, 0.0, 0.0]) + 1e-4 * torch.randn(8, 4, generator=g)  # Taylor branch
    quats[72:80, 0] = -quats[72:80, 0].abs()  # cos_theta < eps branch (covariance_utils.py:42-49)
    scales = torch.exp(torch.randn(512, 3, generator=g) * 0.6 + np.log(0.008))
    covs = cu.compute_cov(quats, scales)
    covs_inv = cu.compute_cov(quats, 1.0 / scales)
    rot = cu.quaternion_to_rotation_matrix(quats)
    np.savez_compressed(OUT / "compute_cov.npz", quats=quats.numpy(), scales=scales.numpy(), covs=covs.numpy(),
                        covs_inv=covs_inv.numpy(), rot=rot.numpy())
    print("compute_cov.npz", covs.shape)


def make_sh2rgb():
    src = (REF / "sim_a_splat/ns_utils/nerfstudio_utils.py").read_text()
    tree = ast.parse(src)
    keep = [n for n in tree.body
            if (isinstance(n, ast.Assign) and getattr(n.targets[0], "id", None) == "C0")
            or (isinstance(n, ast.FunctionDef) and n.name == "SH2RGB")]
    ns: dict = {}
    exec(compile(ast.Module(body=keep, type_ignores=[]), "nerfstudio_utils.py", "exec"), ns)
    g = torch.Generator().manual_seed(11)
    dc = torch.randn(256, 3, generator=g) * 1.5
    np.savez_compressed(OUT / "sh2rgb.npz", C0=np.float64(ns["C0"]), dc=dc.numpy(), rgb=ns["SH2RGB"](dc).numpy())
    print("sh2rgb.npz C0 =", ns["C0"])


class _Handle:
    def __init__(self):
        self.wxyz = None
        self.position = None


class _Quat:
    def __init__(self, w):
        self._w = np.asarray(w, dtype=np.float64)

    def wxyz(self):
        return self._w


class _Rot:
    def ToQuaternion(self):
        return _Quat([1.0, 0.0, 0.0, 0.0])


class _Weld:
    def __init__(self, t):
        self._t = np.asarray(t, dtype=np.float64)

    def translation(self):
        return self._t

    def rotation(self):
        return _Rot()


def _standin_tf():
    """scipy-only stand-in for viser.transforms (tools/viser_tf_scipy.py) -- NOT this repo's transforms.py."""
    sys.path.insert(0, str(REPO / "tools"))
    import viser_tf_scipy

    return viser_tf_scipy


def _import_reference_splat_handler():
    mytf = _standin_tf()

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    viser = stub("viser", ViserServer=object)
    viser.transforms = mytf
    sys.modules["viser.transforms"] = mytf
    stub("drake", lcmt_viewer_draw=object, lcmt_viewer_load_robot=object)
    stub("urchin", URDF=object)
    stub("open3d")
    stub("trimesh")
    stub("pydrake")
    stub("pydrake.all", RigidTransform=lambda: _Weld([0.0, 0.0, 0.0]))
    stub("sim_a_splat")
    stub("sim_a_splat.splat")
    stub("sim_a_splat.splat.splat_utils", GSplatLoader=object)
    stub("sim_a_splat.messaging")
    stub("sim_a_splat.messaging.link", Link=object)
    return _load_by_path("ref_splat_handler", REF / "sim_a_splat/splat/splat_handler.py")


def make_draw_handler():
    mytf = _standin_tf()

    sh = _import_reference_splat_handler()
    rng = np.random.default_rng(2024)
    out = {}
    mask_dirs = {
        "xarm6-1": REF / "assets/robots-scene-v2/masks/xarm6-1",
        "xarm6-2": REF / "assets/robots-scene-v2/masks/xarm6-2",
        "xarm6-robot2": REF / "assets/robots-scene-v2/masks/xarm6-robot2",
        "divar113vhw": REF / "assets/divar113vhw/masks/divar113vhw",
    }
    cases = []
    for name, d in mask_dirs.items():
        h = object.__new__(sh.SplatHandler)
        h.masks_dir = str(d)
        h._load_saved_masks()  # reference code: ICP factorisation + asserts
        out[f"{name}/icp"] = np.load(d / "icp_transformation.npy")
        out[f"{name}/scale_factor"] = np.float64(h.scale_factor)
        out[f"{name}/Ri"] = np.asarray(h.Ri)
        out[f"{name}/ti"] = np.asarray(h.ti)
        L = len(h.link_masks_dict_saved)
        nb = L + 1  # one more robot body than masks, like the xarm push gripper
        for case in range(3):
            weld = rng.uniform(-0.2, 0.2, size=3) if case else np.zeros(3)
            h.sim_robot_weld_frame_transform = _Weld(weld)
            h.rbt_idx, h.blk_idx = 3, 2
            fk = [mytf.SE3.from_rotation_and_translation(mytf.SO3.exp(rng.uniform(-2, 2, 3)), rng.uniform(-0.5, 0.5, 3))
                  for _ in range(nb)]
            h.fk_tf = fk
            h.mesh_frame_handles = [_Handle() for _ in range(nb)]
            h.splat_links_handler = [_Handle() for _ in range(L)]
            h.task_mesh_frame_handle = _Handle()
            # message: T-block first (robot_num 2), then robot bodies (robot_num 3), one foreign body
            B = nb + 2
            msg = types.SimpleNamespace()
            msg.num_links = B
            msg.robot_num = [2] + [3] * nb + [7]
            msg.link_name = ["plant::tblock_paper::tblock_paper"] + [f"plant::robot::link{i}" for i in range(nb)] + ["plant::other::x"]
            q = rng.normal(size=(B, 4)) * rng.uniform(0.5, 2.0, size=(B, 1))  # deliberately un-normalised
            p = rng.uniform(-0.6, 0.6, size=(B, 3))
            msg.quaternion = [q[i] for i in range(B)]
            msg.position = [p[i] for i in range(B)]
            h.draw_handler(msg)  # reference code
            nsplat = min(L, 7)  # `local_idx < 7` (splat_handler.py:282)
            key = f"{name}/case{case}"
            out[f"{key}/weld"] = weld
            out[f"{key}/robot_num"] = np.asarray(msg.robot_num)
            out[f"{key}/quaternion"] = q
            out[f"{key}/position"] = p
            out[f"{key}/fk_wxyz_xyz"] = np.stack([f.wxyz_xyz for f in fk])
            out[f"{key}/splat_wxyz"] = np.stack([h.splat_links_handler[i].wxyz for i in range(nsplat)])
            out[f"{key}/splat_position"] = np.stack([h.splat_links_handler[i].position for i in range(nsplat)])
            out[f"{key}/mesh_wxyz"] = np.stack([m.wxyz for m in h.mesh_frame_handles])
            out[f"{key}/mesh_position"] = np.stack([m.position for m in h.mesh_frame_handles])
            out[f"{key}/task_wxyz"] = np.asarray(h.task_mesh_frame_handle.wxyz)
            out[f"{key}/task_position"] = np.asarray(h.task_mesh_frame_handle.position)
            local = mytf.SE3(np.concatenate(([1.0, 0, 0, 0], [-0.1, 0.0, 0.033])))  # tests/test_splat_wrapper.py:92-98
            aw, ap = h.get_attached_frame("robot::link3", local, msg)  # reference code
            out[f"{key}/attached_wxyz"] = np.asarray(aw)
            out[f"{key}/attached_position"] = np.asarray(ap)
            cases.append(key)
    out["cases"] = np.array(cases)
    np.savez_compressed(OUT / "draw_handler.npz", **out)
    print("draw_handler.npz", len(cases), "cases")


def _read_pcd_xyz(path: Path) -> np.ndarray:
    raw = path.read_bytes()
    head_end = raw.index(b"DATA binary\n") + len(b"DATA binary\n")
    header = raw[:head_end].decode()
    n = int([ln for ln in header.splitlines() if ln.startswith("POINTS")][0].split()[1])
    arr = np.frombuffer(raw[head_end:head_end + n * 24], dtype=np.float32).reshape(n, 6)
    return arr[:, :3].copy()


def make_assets():
    out = {}
    for scene, obj in (("robots-scene-v2", "xarm6-1"), ("divar113vhw", "divar113vhw")):
        d = REF / "assets" / scene / "masks" / obj
        masks = np.load(d / "link_masks_global_dict.npy", allow_pickle=True).item()
        L = len(masks)
        stack = np.stack([masks[f"link{i}"] for i in range(L)])
        out[f"{scene}/N"] = np.int64(stack.shape[1])
        out[f"{scene}/masks_packed"] = np.packbits(stack, axis=1)
        out[f"{scene}/icp"] = np.load(d / "icp_transformation.npy")
        jc = d / "joint_config.npy"
        out[f"{scene}/joint_config"] = np.load(jc) if jc.exists() else np.zeros(0)
        pts = _read_pcd_xyz(d / "point_cloud.pcd")
        sel = np.random.default_rng(5).choice(pts.shape[0], size=8192, replace=False)
        out[f"{scene}/surface_points"] = pts[np.sort(sel)]
        tj = json.loads((REF / "assets" / scene / "transforms.json").read_text())
        out[f"{scene}/intrinsics"] = np.array([tj["w"], tj["h"], tj["fl_x"], tj["fl_y"], tj["cx"], tj["cy"]], dtype=np.float64)
        frames = tj["frames"][:: max(1, len(tj["frames"]) // 32)][:32]
        out[f"{scene}/train_c2w"] = np.stack([np.asarray(f["transform_matrix"], dtype=np.float64) for f in frames])
        sp = sorted((REF / "assets" / scene / "splatfacto").glob("*/dataparser_transforms.json"))[0]
        dp = json.loads(sp.read_text())
        out[f"{scene}/dataparser_transform"] = np.asarray(dp["transform"], dtype=np.float64)
        out[f"{scene}/dataparser_scale"] = np.float64(dp["scale"])
    # camera dict every reference script uses (tests/test_splat_wrapper.py:77-101)
    out["script_camera0_wxyz_xyz"] = np.array([-0.41946813, 0.89955231, -0.11045113, 0.05150421, -0.15, -0.3, -0.05])
    out["script_camera1_wxyz_xyz"] = np.array([1.0, 0, 0, 0, -0.1, 0, 0.033])
    out["script_render_size"] = np.array([240, 320])
    np.savez_compressed(OUT / "assets.npz", **out)
    print("assets.npz", {k: v.shape for k, v in out.items() if hasattr(v, "shape") and v.ndim > 0})


def make_urdf_fk():
    """Visual-mesh forward kinematics of the reference's REAL robot descriptions (xarm6 with push gripper, divar113vhw)
    by an implementation that shares nothing with sim_a_splat_b200/urdf_fk.py: ElementTree + scipy Rotation
    (URDF rpy = Rotation.from_euler('xyz', rpy) in scipy's extrinsic convention; joint motion = from_rotvec(axis q)).
    The fixture holds a STRIPPED copy of each URDF (links, visual origins, joint tree: the numbers the FK needs, none of
    the inertial / collision / material blocks) so that the test can run the product parser without /root/reference,
    plus the transforms at the capture-time configuration (zeros for the xarm, match_splat.py:67's [0.2, 3.0, 3.14, 0, 0]
    for divar113vhw, whose joint_config.npy is absent from the assets) and at a random configuration."""
    import xml.etree.ElementTree as ET

    from scipy.spatial.transform import Rotation as R

    def H(el):
        T = np.eye(4)
        if el is not None:
            T[:3, :3] = R.from_euler("xyz", [float(v) for v in el.get("rpy", "0 0 0").split()]).as_matrix()
            T[:3, 3] = [float(v) for v in el.get("xyz", "0 0 0").split()]
        return T

    out = {}
    rng = np.random.default_rng(77)
    for name, path, q0 in (("xarm6", REF / "robot_description/xarm_description/xarm6/urdf/xarm6_with_push_gripper.urdf", None),
                           ("divar113vhw", REF / "robot_description/divar113vhw/urdf/divar113vhw.urdf", [0.2, 3.0, 3.14, 0, 0])):
        root = ET.parse(path).getroot()
        links = [l.get("name") for l in root.findall("link")]
        joints = root.findall("joint")
        act = [j.get("name") for j in joints if j.get("type") in ("revolute", "continuous", "prismatic") and j.find("mimic") is None]
        # stripped URDF
        lines = [f'<robot name="{name}">']
        for l in root.findall("link"):
            lines.append(f'  <link name="{l.get("name")}">')
            for v in l.findall("visual"):
                o, g = v.find("origin"), v.find("geometry")
                mesh = g.find("mesh") if g is not None else None
                lines.append("    <visual>")
                if o is not None:
                    lines.append(f'      <origin xyz="{o.get("xyz", "0 0 0")}" rpy="{o.get("rpy", "0 0 0")}"/>')
                lines.append(f'      <geometry>{"<mesh filename=" + chr(34) + mesh.get("filename") + chr(34) + "/>" if mesh is not None else "<box size=" + chr(34) + "1 1 1" + chr(34) + "/>"}</geometry>')
                lines.append("    </visual>")
            lines.append("  </link>")
        for j in joints:
            o, ax, mim = j.find("origin"), j.find("axis"), j.find("mimic")
            lines.append(f'  <joint name="{j.get("name")}" type="{j.get("type")}">')
            lines.append(f'    <parent link="{j.find("parent").get("link")}"/><child link="{j.find("child").get("link")}"/>')
            if o is not None:
                lines.append(f'    <origin xyz="{o.get("xyz", "0 0 0")}" rpy="{o.get("rpy", "0 0 0")}"/>')
            if ax is not None:
                lines.append(f'    <axis xyz="{ax.get("xyz")}"/>')
            if mim is not None:
                lines.append(f'    <mimic joint="{mim.get("joint")}" multiplier="{mim.get("multiplier", "1")}" offset="{mim.get("offset", "0")}"/>')
            lines.append("  </joint>")
        lines.append("</robot>")
        out[f"{name}/urdf"] = np.array("\n".join(lines))
        out[f"{name}/actuated"] = np.array(act)

        def fk(q):
            cfg = dict(zip(act, q))
            children = {j.find("child").get("link") for j in joints}
            T = {l: np.eye(4) for l in links if l not in children}
            pend = list(joints)
            while pend:
                for j in list(pend):
                    par = j.find("parent").get("link")
                    if par not in T:
                        continue
                    qj = cfg.get(j.get("name"), 0.0)
                    mim = j.find("mimic")
                    if mim is not None:
                        qj = cfg.get(mim.get("joint"), 0.0) * float(mim.get("multiplier", 1)) + float(mim.get("offset", 0))
                    ax = j.find("axis")
                    a = np.array([float(v) for v in ax.get("xyz").split()]) if ax is not None else np.array([1.0, 0, 0])
                    a = a / (np.linalg.norm(a) or 1.0)
                    M = np.eye(4)
                    if j.get("type") in ("revolute", "continuous"):
                        M[:3, :3] = R.from_rotvec(a * qj).as_matrix()
                    elif j.get("type") == "prismatic":
                        M[:3, 3] = a * qj
                    T[j.find("child").get("link")] = T[par] @ H(j.find("origin")) @ M
                    pend.remove(j)
            vis = []
            for l in root.findall("link"):
                for v in l.findall("visual"):
                    g = v.find("geometry")
                    if g is not None and g.find("mesh") is not None:
                        vis.append(T[l.get("name")] @ H(v.find("origin")))
            return np.stack(vis)

        q_cap = np.zeros(len(act)) if q0 is None else np.asarray(q0, dtype=np.float64)[: len(act)]
        q_rnd = rng.uniform(-1.5, 1.5, len(act))
        out[f"{name}/q_capture"], out[f"{name}/q_random"] = q_cap, q_rnd
        out[f"{name}/fk_capture"], out[f"{name}/fk_random"] = fk(q_cap), fk(q_rnd)
    np.savez_compressed(OUT / "urdf_fk.npz", **out)
    print("urdf_fk.npz", {k: v.shape for k, v in out.items() if getattr(v, "ndim", 0) > 0})


if __name__ == "__main__":
    OUT.mkdir(parents=True, exist_ok=True)
    make_compute_cov()
    make_sh2rgb()
    make_draw_handler()
    make_assets()
    make_urdf_fk()
