"""Synthetic on-disk assets + a fake physics env with the attribute / message surface the reference's
wrappers expect (ManipulatorSimEnv: /root/reference/sim_a_splat/env/manipulator/manipulator_env.py:377-390
for the draw message layout), so the compat classes can be exercised without Drake, viser or the
git-LFS checkpoints."""
from __future__ import annotations

import types
from pathlib import Path

import numpy as np
import torch

from sim_a_splat_b200 import scene as S
from sim_a_splat_b200 import transforms as tf
from sim_a_splat_b200.urdf_fk import SimpleURDF

URDF = """<?xml version="1.0"?>
<robot name="toyarm">
  <material name="Blue"><color rgba="0.2 0.4 0.9 1.0"/></material>
  <link name="world"/>
  <joint name="world_joint" type="fixed"><parent link="world"/><child link="link_base"/><origin rpy="0 0 0" xyz="0 0 0"/></joint>
  <link name="link_base"><visual><geometry><mesh filename="package://toyarm/meshes/base.obj"/></geometry><origin rpy="0 0 0" xyz="0 0 0"/><material name="Blue"/></visual></link>
  <joint name="joint1" type="revolute"><parent link="link_base"/><child link="link1"/><origin rpy="0 0 0" xyz="0 0 0.12"/><axis xyz="0 0 1"/><limit lower="-3" upper="3" effort="1" velocity="1"/></joint>
  <link name="link1"><visual><geometry><mesh filename="package://toyarm/meshes/l1.obj"/></geometry><origin rpy="0 0 0" xyz="0 0 0"/><material name="Blue"/></visual></link>
  <joint name="joint2" type="revolute"><parent link="link1"/><child link="link2"/><origin rpy="-1.5708 0 0" xyz="0 0 0.1"/><axis xyz="0 0 1"/><limit lower="-3" upper="3" effort="1" velocity="1"/></joint>
  <link name="link2"><visual><geometry><mesh filename="package://toyarm/meshes/l2.obj"/></geometry><origin rpy="0 0 0.3" xyz="0.01 0 0"/><material name="Blue"/></visual></link>
  <joint name="joint3" type="prismatic"><parent link="link2"/><child link="gripper"/><origin rpy="0 0 0" xyz="0.15 0 0"/><axis xyz="1 0 0"/><limit lower="0" upper="0.1" effort="1" velocity="1"/></joint>
  <link name="gripper"><visual><geometry><mesh filename="package://toyarm/meshes/g.obj"/></geometry><origin rpy="0 0 0" xyz="0 0 0"/><material name="Blue"/></visual></link>
</robot>
"""
TBLOCK_OBJ = """# T-block, metres (two boxes: bar 0.2 x 0.05, stem 0.05 x 0.15, height 0.04)
v -0.1 -0.025 0.0
v 0.1 -0.025 0.0
v 0.1 0.025 0.0
v -0.1 0.025 0.0
v -0.1 -0.025 0.04
v 0.1 -0.025 0.04
v 0.1 0.025 0.04
v -0.1 0.025 0.04
v -0.025 -0.175 0.0
v 0.025 -0.175 0.0
v 0.025 -0.025 0.0
v -0.025 -0.025 0.0
v -0.025 -0.175 0.04
v 0.025 -0.175 0.04
v 0.025 -0.025 0.04
v -0.025 -0.025 0.04
f 1//1 2//1 3//1 4//1
f 5//1 6//1 7//1 8//1
f 1//1 2//1 6//1 5//1
f 2//1 3//1 7//1 6//1
f 3//1 4//1 8//1 7//1
f 4//1 1//1 5//1 8//1
f 9//1 10//1 11//1 12//1
f 13//1 14//1 15//1 16//1
f 9//1 10//1 14//1 13//1
f 10//1 11//1 15//1 14//1
f 12//1 9//1 13//1 16//1
"""
N_LINKS = 3  # masks for link0..link2; the 4th body (gripper) has a mesh pose but no splat group, like the xarm push gripper


def make_assets(root: Path, assets: dict, n_gaussians: int = 6000, sh_degree: int = 2, seed: int = 5) -> dict:
    """Writes masks/, splatfacto/ (a checkpoint in the nerfstudio key layout) and robot_description/ under root."""
    root = Path(root)
    obj, cfg = "toyarm-1", "2024-01-01_000000"
    icp = assets["robots-scene-v2/icp"]
    urdf_dir = root / "robot_description" / "toyarm" / "urdf"
    urdf_dir.mkdir(parents=True)
    (urdf_dir / "toyarm.urdf").write_text(URDF)
    joint_config = np.array([0.2, -0.4, 0.03])
    u = SimpleURDF.load(str(urdf_dir / "toyarm.urdf"))
    fk = u.visual_mesh_fk(dict(zip(u.actuated_joint_names, joint_config)))
    masks = S.scatter_masks(n_gaussians, (400, 300, 200), seed + 1)
    rng = np.random.default_rng(seed)
    # link Gaussians sit near their link's capture-time pose, mapped into the splat frame by the ICP transform
    link_points = [S.icp_apply(icp, (T[:3, :3] @ rng.normal(0, 0.02, size=(256, 3)).T).T + T[:3, 3]) for T in fk[:N_LINKS]]
    means, quats, scales, opac, sh = S.synthetic_params(n_gaussians, sh_degree, seed, masks, link_points)
    mdir = root / "masks" / obj
    mdir.mkdir(parents=True)
    np.save(mdir / "link_masks_global_dict.npy", {f"link{i}": masks[i] for i in range(N_LINKS)}, allow_pickle=True)
    np.save(mdir / "icp_transformation.npy", icp)
    np.save(mdir / "joint_config.npy", joint_config)
    cdir = root / "splatfacto" / cfg / "nerfstudio_models"
    cdir.mkdir(parents=True)
    pre = "_model.gauss_params."
    logit = lambda p: np.log(p / (1 - p))  # noqa: E731
    state = {pre + "means": torch.from_numpy(means), pre + "quats": torch.from_numpy(quats),
             pre + "scales": torch.from_numpy(np.log(scales)), pre + "opacities": torch.from_numpy(logit(opac.astype(np.float64)).astype(np.float32))[:, None],
             pre + "features_dc": torch.from_numpy(sh[:, 0]), pre + "features_rest": torch.from_numpy(sh[:, 1:])}
    torch.save({"pipeline": state, "step": 1}, cdir / "step-000000001.ckpt")
    tdir = root / "tblock_paper"
    tdir.mkdir(parents=True)
    (tdir / "tblock_paper.obj").write_text(TBLOCK_OBJ)
    return dict(root=str(root), obj=obj, cfg=cfg, package_path=str(root / "robot_description"), package_name="toyarm/", urdf_name="toyarm.urdf",
                masks=masks, joint_config=joint_config, n=n_gaussians, task_assets_path=str(tdir), task_assets_name="tblock_paper.obj")


class FakeManipulatorEnv:
    """Physics stand-in: joint angles random-walk; body poses come from the URDF FK; the draw message has the
    lcmt_viewer_draw field layout (robot_num 2 = T-block first, then robot_num 3 = robot bodies)."""

    def __init__(self, info: dict, seed: int = 0):
        self.package_path, self.package_name, self.urdf_name = info["package_path"], info["package_name"], info["urdf_name"]
        self.weld_frame_transform = None
        self.visualize_robot_flag = False
        self.urdf = SimpleURDF.load(f"{self.package_path}/{self.package_name}urdf/{self.urdf_name}")
        self.rng = np.random.default_rng(seed)
        self.q = np.zeros(len(self.urdf.actuated_joint_names))
        self.block = np.array([0.3, 0.1, 0.0, 0.0])

    @property
    def unwrapped(self):
        return self

    def reset(self, seed=None, reset_to_state=None):
        if seed is not None:
            self.rng = np.random.default_rng(seed)
        self.q = self.rng.uniform(-0.5, 0.5, size=self.q.shape) * np.array([1, 1, 0.1])
        self.q[2] = abs(self.q[2])
        return self._get_obs()

    def step(self, action):
        self.q = self.q + np.asarray(action, dtype=np.float64)
        return self._get_obs(), 0.0, False, False, {}

    def _get_obs(self):
        return {"joint_pos": self.q.copy(), "joint_vel": np.zeros_like(self.q)}

    def _generate_draw_msg(self):
        T = self.urdf.link_fk(dict(zip(self.urdf.actuated_joint_names, self.q)))
        bodies = [v.link for v in self.urdf.mesh_visuals()]
        msg = types.SimpleNamespace()
        names, nums, pos, quat = ["plant::tblock_paper::tblock_paper"], [2], [self.block[:3].copy()], [tf.SO3.from_z_radians(self.block[3]).wxyz]
        for b in bodies:
            names.append(f"plant::toyarm::{b}")
            nums.append(3)
            pos.append(T[b][:3, 3].copy())
            quat.append(tf.SO3.from_matrix(T[b][:3, :3]).wxyz * 1.7)  # deliberately un-normalised, like raw LCM floats can be
        msg.num_links, msg.link_name, msg.robot_num, msg.position, msg.quaternion = len(names), names, nums, pos, quat
        return msg

    def render(self):
        return None

    def close(self):
        pass


def camera_setup(render_size=(64, 80)):
    """The camera dict every reference script uses (tests/test_splat_wrapper.py:77-101), smaller images."""
    return {
        0: {"link_name": "world", "type": "viewport", "render_size": list(render_size),
            "local_frame": tf.SE3(np.concatenate((np.array([-0.41946813, 0.89955231, -0.11045113, 0.05150421]), np.array([-0.15, -0.3, -0.05]))))},
        1: {"link_name": "toyarm::gripper", "type": "moving", "render_size": list(render_size),
            "local_frame": tf.SE3(np.concatenate((np.array([1.0, 0, 0, 0]), np.array([-0.1, 0, 0.033]))))},
    }
