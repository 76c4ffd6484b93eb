"""GaussianSplat with the reference's render contract
(/root/reference/sim_a_splat/ns_utils/nerfstudio_utils.py:50-177): `render(pose)` takes a camera-to-world
pose in nerfstudio's convention (OpenGL camera axes, [3|4, 4]) and returns the splatfacto output dict
`{"rgb": [H,W,3] in [0,1], "depth": [H,W,1] expected depth, "accumulation": [H,W,1], "background": [3]}`
rendered at the intrinsics of training camera 0 (:125-138, :166-172).  This is the only gsplat call site of
the reference; here the frame comes from libsas_b200 with `nerfstudio_post=1`.

The nerfstudio pipeline (`eval_setup`, datamanager, dataset images) is not reproduced: the constructor
takes the Gaussians (a SplatScene, a GSplatLoader, or a splatfacto output directory) and the training
cameras (`transforms.json`-style intrinsics + camera_to_worlds).  Point-cloud export and CLIP queries
(:179-472) are out of scope.
"""
from __future__ import annotations

from pathlib import Path
from typing import Optional, Union

import numpy as np
import torch

from ..renderer import NS_BACKGROUND, SplatRenderer
from ..scene import C0, SplatScene, c2w_opengl_to_viewmat, canonicalize


def SH2RGB(sh):
    return sh * C0 + 0.5


class GaussianSplat:
    def __init__(self, config_path: Union[Path, SplatScene, object], res_factor=None, test_mode="inference", dataset_mode="test",
                 device: Union[torch.device, str, int] = 0, intrinsics=None, camera_to_worlds=None, background=NS_BACKGROUND):
        """intrinsics = (H, W, K[3,3]) of training camera 0; camera_to_worlds [n,3,4] (nerfstudio convention)."""
        self.config_path, self.res_factor = config_path, res_factor
        self.device = torch.device(f"cuda:{device}" if isinstance(device, int) else device)
        if isinstance(config_path, SplatScene):
            self.scene = config_path
        else:
            from ..splat.splat_utils import GSplatLoader

            gs = config_path if hasattr(config_path, "means") else GSplatLoader(Path(config_path), "cpu")
            n = gs.means.shape[0]
            self.scene = canonicalize(gs.means.cpu().numpy(), gs.rots.cpu().numpy(), gs.scales.cpu().numpy(),
                                      gs.opacities.cpu().numpy().reshape(n), gs.sh.cpu().numpy(), None)
        if intrinsics is None:
            raise ValueError("intrinsics=(H, W, K) of training camera 0 is required (the reference reads it from the dataset)")
        H, W, K = intrinsics
        K = np.asarray(K, dtype=np.float64).reshape(3, 3).copy()
        if res_factor is not None:  # Cameras.rescale_output_resolution (:99-100)
            K[:2] *= res_factor
            H, W = int(H * res_factor), int(W * res_factor)
        self.H, self.W, self.K = int(H), int(W), K
        self.camera_to_worlds = None if camera_to_worlds is None else torch.as_tensor(np.asarray(camera_to_worlds), dtype=torch.float32)
        self.background = tuple(float(b) for b in background)
        self._renderer: Optional[SplatRenderer] = None

    def get_poses(self):
        return self.camera_to_worlds

    def get_camera_intrinsics(self):
        return self.H, self.W, torch.as_tensor(self.K, dtype=torch.float32)

    def render(self, pose, compute_semantics: Optional[bool] = False, debug_mode=False) -> dict:
        c2w = pose.detach().cpu().numpy() if isinstance(pose, torch.Tensor) else np.asarray(pose)
        if self._renderer is None:
            dev = self.device.index or 0
            self._renderer = SplatRenderer(self.scene, device=dev, max_images=1, max_envs=1, max_width=self.W, max_height=self.H)
        r = self._renderer
        r.set_cameras(c2w_opengl_to_viewmat(c2w[:3])[None], self.K, self.W, self.H)
        out = r.render(("rgb", "depth", "alpha"), background=self.background, nerfstudio_post=True)
        return {"rgb": out["rgb"][0], "depth": out["depth"][0][..., None], "accumulation": out["alpha"][0][..., None],
                "background": torch.tensor(self.background, device=out["rgb"].device)}

    def generate_RGBD_point_cloud(self, pose, max_depth: float = 1.0, return_pcd: bool = False):
        """Render at `pose` and unproject every pixel with depth < max_depth through the pinhole model of the
        reference (nerfstudio_utils.py:375-472; the arithmetic is :441-452): returns (points [M,3] in the CAMERA
        frame with OpenCV axes, colours [M,3], depth image [H,W], rgb image [H,W,3]).  The open3d point-cloud
        object, outlier removal and world transform of the reference are not reproduced."""
        out = self.render(pose)
        depth = out["depth"][..., 0]
        xyz, valid = self._renderer.unproject_depth(depth[None], self.K, max_depth)
        return xyz[0][valid[0]], out["rgb"][valid[0]], depth, out["rgb"]
