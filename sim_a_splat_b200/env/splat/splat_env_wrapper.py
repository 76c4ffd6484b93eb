"""SplatEnvWrapper with the reference's constructor, methods and observation layout
(/root/reference/sim_a_splat/env/splat/splat_env_wrapper.py:13-163): a gym.Wrapper that glues a physics env
(anything exposing `_generate_draw_msg()`, `_get_obs()`, `package_path`, `package_name`, `urdf_name`,
`weld_frame_transform`, `visualize_robot_flag`) to a SplatHandler; `step(action, noobs=False)` adds
`camera_{i}` = uint8 `(3,H,W)` per configured camera (:132-138).

Differences, deliberate: no browser client to wait for (the headless client exists at once);
`step` stores the new draw message in `self.draw_msg`, so link-attached ("moving") cameras follow the robot
-- the reference keeps the reset-time message (:99, :123, :145); pass `reference_quirks=True` to reproduce
that; all cameras of a step are rendered in one renderer call instead of one blocking RPC each.
"""
from __future__ import annotations

import logging
from typing import Optional

import numpy as np

from ... import transforms as tf
from ...splat.splat_handler import SplatHandler

try:  # gymnasium is not installed in the build image; the wrapper only needs .env / .unwrapped
    import gymnasium as gym

    _Base = gym.Wrapper
except Exception:  # pragma: no cover - exercised offline
    class _Base:  # minimal gym.Wrapper surface
        def __init__(self, env):
            self.env = env

        @property
        def unwrapped(self):
            return getattr(self.env, "unwrapped", self.env)


class SplatEnvWrapper(_Base):
    def __init__(self, env, splat_assets_path: str, match_object_name: str, splat_config_name: str,
                 task_assets_path: Optional[str] = None, task_assets_name: Optional[str] = None, reference_quirks: bool = False):
        super().__init__(env)
        self.reference_quirks = reference_quirks
        self.ch = self._setup_splats(splat_assets_path=splat_assets_path, match_object_name=match_object_name,
                                     splat_config_name=splat_config_name, task_assets_path=task_assets_path,
                                     task_assets_name=task_assets_name)

    def _configure_cameras(self, camera_setup_info: dict):
        viewport = {k: v for k, v in camera_setup_info.items() if v.get("type") == "viewport"}
        self.moving_cameras_info = {k: v for k, v in camera_setup_info.items() if v.get("type") == "moving"}
        fixed = {k: v for k, v in camera_setup_info.items() if v.get("type") in ("viewport", "static")}
        self.fixed_cam_poses = [info["local_frame"] for info in fixed.values()]
        self.render_cam_keys = list(self.moving_cameras_info.keys()) + list(fixed.keys())
        first = viewport[list(viewport.keys())[0]]["local_frame"]
        self.ch.camera.position = first.translation()
        self.ch.camera.wxyz = first.rotation().wxyz
        self.camera_setup_info = camera_setup_info

    def _setup_splats(self, splat_assets_path, match_object_name, splat_config_name, task_assets_path, task_assets_name, wait_steps=50):
        u = self.unwrapped
        self.splat_handler = SplatHandler(splat_assets_path, match_object_name, splat_config_name, u.package_path, u.package_name,
                                          u.urdf_name, task_assets_path=task_assets_path, task_assets_name=task_assets_name,
                                          sim_robot_weld_frame_transform=getattr(u, "weld_frame_transform", None))
        clients = self.splat_handler.server.get_clients()
        if len(clients) == 0:
            raise RuntimeError("No clients connected after waiting")
        logging.info(f"client: {clients}")
        return clients[0]

    def reset(self, seed: Optional[int] = None, reset_to_state=None):
        self.unwrapped.reset(seed=seed, reset_to_state=reset_to_state)
        self.draw_msg = self.unwrapped._generate_draw_msg()
        self.splat_handler.draw_handler(self.draw_msg)
        if getattr(self.unwrapped, "visualize_robot_flag", False):
            self.env.render()
        return self.unwrapped._get_obs()

    def get_moving_camera_poses(self, msg):
        poses = []
        try:
            for info in self.moving_cameras_info.values():
                wxyz, xyz = self.splat_handler.get_attached_frame(info["link_name"], info["local_frame"], msg)
                poses.append(tf.SE3(np.concatenate((wxyz, xyz))))
        except AttributeError as e:
            logging.error(f"Error getting moving camera poses: {e}. Have you configured the cameras with _configure_cameras?")
        return poses

    def step(self, action, noobs=False):
        obs_in, reward, terminated, truncated, info_in = self.env.step(action)
        draw_msg = self.unwrapped._generate_draw_msg()
        if not self.reference_quirks:
            self.draw_msg = draw_msg
        self.splat_handler.draw_handler(draw_msg)
        observation = None
        if not noobs:
            if getattr(self.unwrapped, "visualize_robot_flag", False):
                self.env.render()
            observation = self._get_obs()
        return observation, reward, terminated, truncated, info_in

    def _get_obs(self):
        obs = self.unwrapped._get_obs()
        img_out = self.render()
        obs.update({f"camera_{ii}": np.moveaxis(img_out[ii], -1, 0) for ii in range(len(img_out))})
        return obs

    def render(self, mode="rgb_array"):
        if hasattr(self.env, "render"):
            self.env.render()
        poses = self.get_moving_camera_poses(self.draw_msg) + self.fixed_cam_poses
        views = []
        for ii, pose in enumerate(poses):
            h, w = self.camera_setup_info[self.render_cam_keys[ii]]["render_size"][:2]
            views.append((int(h), int(w), pose.rotation().wxyz, pose.translation()))
        return self.splat_handler.server.render_views(views, self.ch.camera.fov)

    def close(self):
        self.splat_handler.server.stop()
        self.unwrapped.close()
