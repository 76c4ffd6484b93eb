"""GSplatLoader with the reference's attribute surface
(/root/reference/sim_a_splat/splat/splat_utils.py:13-89): `means, rots, scales, covs, covs_inv, colors,
opacities` as torch tensors with exp / sigmoid applied (:33-45, :77-85), plus `sh` ([N,K,3], features_dc ++
features_rest) which the gsplat-semantics render needs and the reference reaches through the model object.

A `str` location is the reference's JSON dump (:51-89).  A `Path` is a nerfstudio splatfacto output
directory (:24-49) -- the reference goes through `eval_setup`; nerfstudio is a 60-package dependency that
is not installable offline, so the checkpoint is read directly: the newest
`nerfstudio_models/step-*.ckpt` is a torch pickle whose `pipeline` state dict holds
`_model.gauss_params.{means,scales,quats,features_dc,features_rest,opacities}` [RECALLED layout].  A
`gaussians.npz` with those six arrays in the directory is accepted as well.
"""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np
import torch

from ..ellipsoids.covariance_utils import compute_cov, quaternion_to_rotation_matrix
from ..scene import C0


def SH2RGB(sh):
    return sh * C0 + 0.5


class GSplatLoader:
    def __init__(self, gsplat_location, device="cpu"):
        self.device = device
        if isinstance(gsplat_location, str):
            self.load_gsplat_from_json(gsplat_location)
        elif isinstance(gsplat_location, Path):
            self.load_gsplat_from_nerfstudio(gsplat_location)
        else:
            raise ValueError("GSplat file must be either a .json or .yml file.")

    # raw (pre-activation) parameters -> the reference's attributes
    def _finish(self, means, quats, log_scales, opacity_logits, features_dc, features_rest=None):
        t = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32, device=self.device)  # noqa: E731
        self.means, self.rots = t(means), t(quats)
        self.scales = torch.exp(t(log_scales))
        self.covs_inv = compute_cov(self.rots, 1.0 / self.scales)
        self.covs = compute_cov(self.rots, self.scales)
        dc = t(features_dc).reshape(self.means.shape[0], 3)
        self.colors = SH2RGB(dc)
        self.opacities = torch.sigmoid(t(opacity_logits)).reshape(-1, 1)
        rest = t(features_rest).reshape(self.means.shape[0], -1, 3) if features_rest is not None and np.size(features_rest) else None
        self.sh = dc[:, None, :] if rest is None else torch.cat([dc[:, None, :], rest], dim=1)

    def load_gsplat_from_nerfstudio(self, gsplat_location: Path):
        d = gsplat_location if gsplat_location.is_dir() else gsplat_location.parent
        npz = d / "gaussians.npz"
        if npz.exists():
            z = np.load(npz)
            p = {k: z[k] for k in z.files}
        else:
            ckpts = sorted((d / "nerfstudio_models").glob("step-*.ckpt"))
            if not ckpts:
                raise FileNotFoundError(f"no nerfstudio_models/step-*.ckpt (or gaussians.npz) under {d}")
            if ckpts[-1].stat().st_size < 4096 and ckpts[-1].read_bytes().startswith(b"version https://git-lfs"):
                raise FileNotFoundError(f"{ckpts[-1]} is a git-LFS pointer, not the checkpoint (fetch the LFS object)")
            state = torch.load(ckpts[-1], map_location="cpu", weights_only=False)["pipeline"]
            pre = "_model.gauss_params."
            p = {k[len(pre):]: v.detach().cpu().numpy() for k, v in state.items() if k.startswith(pre)}
        self._finish(p["means"], p["quats"], p["scales"], p["opacities"], p["features_dc"], p.get("features_rest"))
        print(f"There are {self.means.shape[0]} Gaussians in the GSplat model")

    def load_gsplat_from_json(self, gsplat_location: str):
        with open(gsplat_location, "r") as f:
            data = json.load(f)
        n = len(data["means"])
        colors = np.asarray(data["colors"], dtype=np.float32).reshape(n, -1)
        self._finish(data["means"], data["rotations"], data["scalings"], data["opacities"], (colors[:, :3] - 0.5) / C0)
        self.colors = torch.as_tensor(colors, dtype=torch.float32, device=self.device)  # the JSON stores RGB, :69

    def rotation_matrices(self):
        return quaternion_to_rotation_matrix(self.rots)


class PointCloudLoader(GSplatLoader):
    def __init__(self, device):
        self.device = device

    def initialize_attributes(self, means):
        self.means = means.to(self.device)


class DummyGSplatLoader(GSplatLoader):
    """Toy examples (splat_utils.py:132-149)."""

    def __init__(self, device):
        self.device = device

    def initialize_attributes(self, means, rots, scales, colors=None):
        self.means, self.rots, self.scales = means.to(self.device), rots.to(self.device), scales.to(self.device)
        self.cov_inv = compute_cov(self.rots, 1 / self.scales)
        self.covs = compute_cov(self.rots, self.scales)
        self.colors = colors.to(self.device) if colors is not None else 0.5 * torch.ones(means.shape[0], 3).to(self.device)
        self.opacities = torch.ones(means.shape[0], 1, device=self.device)
        self.sh = ((self.colors - 0.5) / C0)[:, None, :]
