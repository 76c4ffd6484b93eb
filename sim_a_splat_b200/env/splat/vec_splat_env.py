"""Batched observation rendering for E parallel envs -- the extension BASELINE.json asks for on top of the
reference's single-env wrapper (the reference renders one camera of one env per browser round trip,
/root/reference/sim_a_splat/env/splat/splat_env_wrapper.py:147-158).

One `VecSplatRenderer.render(draw_msgs)` call = one pass of the hot path over E x C images: the per-env link
poses are composed on the device from the draw messages' body poses (the draw_handler algebra,
`sas_set_link_poses_from_bodies`), fixed cameras are shared
views, link-attached cameras get one view per env.  Observations keep the reference layout: `camera_{i}`
= uint8 [E,3,H,W] (splat_env_wrapper.py:132-138).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np

from ...renderer import PipelinedSplatRenderer, SplatRenderer
from ...scene import c2w_opencv_to_viewmat
from ...splat.splat_handler import SplatHandler


class VecSplatRenderer:
    def __init__(self, handler: SplatHandler, camera_setup_info: dict, n_envs: int, fov: float = np.radians(75.0),
                 background=(1.0, 1.0, 1.0), device: int = 0, max_isects: int = 0, env_range: Optional[range] = None,
                 pipeline_depth: int = 0):
        """env_range: the envs this process renders (multi-GPU sharding, see sim_a_splat_b200/dist.py).
        pipeline_depth > 0: render_async / render_wait keep that many steps in flight on as many renderer handles
        (PipelinedSplatRenderer); render() is then not available."""
        self.handler, self.info, self.fov, self.background = handler, camera_setup_info, fov, background
        self.envs = list(env_range) if env_range is not None else list(range(n_envs))
        self.moving = {k: v for k, v in camera_setup_info.items() if v.get("type") == "moving"}
        self.fixed = {k: v for k, v in camera_setup_info.items() if v.get("type") in ("viewport", "static")}
        self.keys = list(self.moving.keys()) + list(self.fixed.keys())  # reference order: moving first (:53-55)
        sizes = {tuple(camera_setup_info[k]["render_size"][:2]) for k in self.keys}
        if len(sizes) != 1:
            raise ValueError("VecSplatRenderer renders one image size per call; group cameras by render_size")
        self.H, self.W = (int(v) for v in sizes.pop())
        scene = handler.server.build_scene()
        self.E, self.C = len(self.envs), len(self.keys)
        kw = dict(device=device, max_images=self.E * self.C, max_envs=self.E, max_width=self.W, max_height=self.H, max_isects=max_isects)
        self.renderer = PipelinedSplatRenderer(scene, depth=pipeline_depth, **kw) if pipeline_depth > 0 else SplatRenderer(scene, **kw)
        fy = self.H / (2.0 * np.tan(fov / 2.0))
        self.K = np.array([[fy, 0, self.W / 2.0], [0, fy, self.H / 2.0], [0, 0, 1.0]])
        if self.renderer.L:
            self.renderer.set_pose_model(**handler.pose_model())  # link poses are composed on the device from body poses
        self._fixed_vm = [c2w_opencv_to_viewmat(v["local_frame"].rotation().wxyz, v["local_frame"].translation()) for v in self.fixed.values()]
        self._static_cameras_set = False

    def _set_step_inputs(self, draw_msgs) -> None:
        """draw_msgs: E draw messages (manipulator_env.py:377-390) -- or, for fixed cameras only, the body poses of all
        envs at once as an [E,L,7] fp64 array / CUDA tensor (quaternion wxyz | position per link body, e.g. from
        urdf_fk.BatchedBodyPoseSource.body_poses): no per-env Python on the step path."""
        r = self.renderer
        if not isinstance(draw_msgs, (list, tuple)):
            if self.moving:
                raise ValueError("batched body poses carry no draw messages: link-attached cameras need render(draw_msgs)")
            if r.L:
                r.set_link_poses_from_bodies(draw_msgs)
            if not self._static_cameras_set:
                vms = self._fixed_vm * self.E
                env = np.repeat(np.arange(self.E, dtype=np.int32), self.C)
                r.set_cameras(np.stack(vms), self.K, self.W, self.H, env)
                self._static_cameras_set = True
            return
        assert len(draw_msgs) == self.E
        if r.L:  # draw_handler's algebra runs on the device over all (env, link) at once (csrc/sas_pose.cu)
            r.set_link_poses_from_bodies(np.stack([self.handler.body_poses_from_msg(m) for m in draw_msgs]))  # [E,L,7] fp64
        if self.moving or not self._static_cameras_set:
            vms, env = [], []
            for e, msg in enumerate(draw_msgs):
                for info in self.moving.values():
                    wxyz, xyz = self.handler.get_attached_frame(info["link_name"], info["local_frame"], msg)
                    vms.append(c2w_opencv_to_viewmat(wxyz, xyz))
                vms += self._fixed_vm
                env += [e] * self.C
            if isinstance(r, PipelinedSplatRenderer):  # moving cameras belong to the step, fixed ones to every handle
                r.set_cameras(np.stack(vms), self.K, self.W, self.H, np.asarray(env, dtype=np.int32), current_only=bool(self.moving))
            else:
                r.set_cameras(np.stack(vms), self.K, self.W, self.H, np.asarray(env, dtype=np.int32))
            self._static_cameras_set = True

    def render(self, draw_msgs: Sequence, outputs=("rgb_u8",), to_host: bool = True) -> Dict[str, np.ndarray]:
        if isinstance(self.renderer, PipelinedSplatRenderer):
            raise RuntimeError("pipeline_depth > 0: use render_async / render_wait")
        self._set_step_inputs(draw_msgs)
        return self._as_obs(self.renderer.render(outputs, background=self.background, to_host=to_host), to_host)

    def render_async(self, draw_msgs: Sequence, outputs=("rgb_u8",)) -> int:
        """Issue the render of this step's observations and return at once (a ticket for render_wait): the transfer
        of step k's observations to the host overlaps the render of step k+1 -- the step_async / step_wait split of
        gymnasium's vector envs.  Issue k+1 before waiting for k; the arrays of k are reused by k+2."""
        self._set_step_inputs(draw_msgs)
        return self.renderer.render_async(outputs, background=self.background)

    def render_wait(self, ticket: Optional[int] = None) -> Dict[str, np.ndarray]:
        return self._as_obs(self.renderer.wait(ticket), True)

    def _as_obs(self, out: dict, to_host: bool) -> Dict[str, np.ndarray]:
        res = {}
        for name, t in out.items():
            a = t.numpy() if to_host else t
            a = a.reshape(self.E, self.C, *a.shape[1:])
            for ci in range(self.C):
                img = a[:, ci]
                if name == "rgb_u8":
                    res[f"camera_{ci}"] = np.moveaxis(img, -1, 1) if to_host else img.permute(0, 3, 1, 2)
                else:
                    res[f"{name}_{ci}"] = img
        return res

    def close(self):
        self.renderer.close()
