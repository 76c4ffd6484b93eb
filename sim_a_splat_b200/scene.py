"""Scene construction for the splat render path: asset ingestion, the canonical
Gaussian order, and the seeded synthetic scenes of SURVEY.md section 8d.

The trained checkpoints of the reference are git-LFS stubs, so benchmark and
parity scenes are synthetic Gaussians with the REAL Gaussian counts, link-mask
structure, ICP matrices and camera metadata of the reference's assets
(tests/golden/assets.npz, produced by tools/make_golden.py).

Canonical order (what the renderer and the oracle both consume): the background
`means[~union(masks)]` in original index order (splat_handler.py:113-119)
followed by the copies of link 0..L-1, each in original index order
(splat_handler.py:121-143; a Gaussian claimed by several masks is duplicated,
exactly as the reference registers one viser splat node per link).
"""
from __future__ import annotations

from dataclasses import dataclass
from pathlib import Path
from typing import Dict, Optional, Sequence

import numpy as np

C0 = 0.28209479177387814  # nerfstudio_utils.py:43


def SH2RGB(sh):
    """nerfstudio_utils.py:46-47"""
    return sh * C0 + 0.5


@dataclass
class SplatScene:
    """Canonical-order Gaussian scene on the host (numpy fp32)."""

    means: np.ndarray      # [N,3]
    quats: np.ndarray      # [N,4] wxyz as stored (not normalised)
    scales: np.ndarray     # [N,3] linear (exp applied once on the host)
    opacities: np.ndarray  # [N]   linear (sigmoid applied once on the host)
    sh: np.ndarray         # [N,K,3], K=(sh_degree+1)^2, [:,0] = features_dc
    link_id: np.ndarray    # [N] int32, -1 = background
    L: int
    orig_index: np.ndarray  # [N] index into the un-partitioned checkpoint order

    @property
    def N(self) -> int:
        return int(self.means.shape[0])

    @property
    def sh_degree(self) -> int:
        return int(round(np.sqrt(self.sh.shape[1]))) - 1

    @property
    def n_background(self) -> int:
        return int(np.sum(self.link_id < 0))

    def as_oracle_dict(self) -> Dict[str, np.ndarray]:
        return dict(means=self.means, quats=self.quats, scales=self.scales, opacities=self.opacities,
                    sh=self.sh, link_id=self.link_id, L=self.L)


def load_reference_assets(path: Optional[str] = None) -> Dict[str, np.ndarray]:
    """tests/golden/assets.npz (real masks / ICP / cameras of the reference's assets)."""
    p = Path(path) if path else Path(__file__).resolve().parents[1] / "tests" / "golden" / "assets.npz"
    z = np.load(p)
    return {k: z[k] for k in z.files}


def unpack_masks(assets: Dict[str, np.ndarray], scene: str) -> np.ndarray:
    """-> bool [L, N] (the `link_masks_global_dict.npy` content, splat_handler.py:62-66)."""
    N = int(assets[f"{scene}/N"])
    return np.unpackbits(assets[f"{scene}/masks_packed"], axis=1)[:, :N].astype(bool)


def canonicalize(means, quats, scales, opacities, sh, masks: Optional[np.ndarray]) -> SplatScene:
    """Partition un-ordered Gaussian parameters into the canonical order."""
    N = means.shape[0]
    if masks is None or len(masks) == 0:
        order = np.arange(N)
        link_id = -np.ones(N, dtype=np.int32)
        L = 0
    else:
        masks = np.asarray(masks, dtype=bool)
        L = masks.shape[0]
        union = masks.any(axis=0)
        parts = [np.nonzero(~union)[0]]
        ids = [-np.ones(parts[0].shape[0], dtype=np.int32)]
        for l in range(L):
            idx = np.nonzero(masks[l])[0]
            parts.append(idx)
            ids.append(np.full(idx.shape[0], l, dtype=np.int32))
        order = np.concatenate(parts)
        link_id = np.concatenate(ids)
    f32 = lambda a: np.ascontiguousarray(np.asarray(a)[order], dtype=np.float32)  # noqa: E731
    return SplatScene(f32(means), f32(quats), f32(scales), f32(np.asarray(opacities).reshape(N)), f32(sh),
                      link_id.astype(np.int32), L, order.astype(np.int64))


def synthetic_params(N: int, sh_degree: int, seed: int, link_masks: Optional[np.ndarray] = None,
                     link_points: Optional[Sequence[np.ndarray]] = None):
    """Seeded Gaussian parameters per SURVEY.md 8d (un-partitioned checkpoint order)."""
    rng = np.random.default_rng(seed)
    n_plane = int(0.6 * N)
    means = np.empty((N, 3), dtype=np.float64)
    # 60 % on planar patches: a table (z ~ -0.3) and two walls; 40 % uniform in the unit cube
    which = rng.integers(0, 3, size=n_plane)
    uv = rng.uniform(-1.0, 1.0, size=(n_plane, 2))
    jitter = rng.normal(0.0, 0.004, size=n_plane)
    pl = np.empty((n_plane, 3))
    pl[which == 0] = np.stack([uv[which == 0, 0], uv[which == 0, 1], -0.3 + jitter[which == 0]], axis=1)
    pl[which == 1] = np.stack([uv[which == 1, 0], 0.9 + jitter[which == 1], uv[which == 1, 1]], axis=1)
    pl[which == 2] = np.stack([-0.9 + jitter[which == 2], uv[which == 2, 0], uv[which == 2, 1]], axis=1)
    means[:n_plane] = pl
    means[n_plane:] = rng.uniform(-1.0, 1.0, size=(N - n_plane, 3))
    perm = rng.permutation(N)
    means = means[perm]
    if link_masks is not None and link_points is not None:
        for l in range(link_masks.shape[0]):
            idx = np.nonzero(link_masks[l])[0]
            pts = link_points[l]
            pick = rng.integers(0, pts.shape[0], size=idx.shape[0])
            d = rng.normal(size=(idx.shape[0], 3))
            d *= (rng.uniform(0.0, 0.015, size=(idx.shape[0], 1)) / np.linalg.norm(d, axis=1, keepdims=True))
            means[idx] = pts[pick] + d  # within 1.5 cm of the link surface (match_splat.py:250)
    log_s = np.clip(rng.normal(np.log(0.008), 0.6, size=(N, 3)), np.log(0.001), np.log(0.08))
    quats = rng.normal(size=(N, 4))
    quats /= np.linalg.norm(quats, axis=1, keepdims=True)
    opac = 0.1 + (0.995 - 0.1) * rng.beta(2.0, 1.2, size=N)
    K = (sh_degree + 1) ** 2
    sh = np.empty((N, K, 3))
    sh[:, 0] = (rng.uniform(0.05, 0.95, size=(N, 3)) - 0.5) / C0
    if K > 1:
        sh[:, 1:] = rng.normal(0.0, 0.04, size=(N, K - 1, 3))
    # host-side activations, once (splat_utils.py:35-36, :43-45): the renderer takes linear values
    scales = np.exp(log_s.astype(np.float32))
    return (means.astype(np.float32), quats.astype(np.float32), scales.astype(np.float32),
            opac.astype(np.float32), sh.astype(np.float32))


def scatter_masks(N: int, counts: Sequence[int], seed: int, overlap_frac: float = 0.1) -> np.ndarray:
    """Synthetic link masks with scattered indices and some multiplicity-2 overlap."""
    rng = np.random.default_rng(seed)
    L = len(counts)
    masks = np.zeros((L, N), dtype=bool)
    for l, c in enumerate(counts):
        masks[l, rng.choice(N, size=min(int(c), N), replace=False)] = True
    for l in range(1, L):
        prev = np.nonzero(masks[l - 1])[0]
        k = int(overlap_frac * min(counts[l], prev.shape[0]))
        if k > 0:
            masks[l, rng.choice(prev, size=k, replace=False)] = True
    return masks


def tblock_points(n: int, seed: int) -> np.ndarray:
    """Surface samples of the PushT T-block (bar 0.2 x 0.05, stem 0.05 x 0.15, height 0.04 m;
    dimensions of assets/tblock_paper/tblock_paper.obj), metric, block frame."""
    rng = np.random.default_rng(seed)
    boxes = [((-0.1, -0.025, 0.0), (0.1, 0.025, 0.04)), ((-0.025, -0.175, 0.0), (0.025, -0.025, 0.04))]
    out = []
    for lo, hi in boxes:
        lo, hi = np.array(lo), np.array(hi)
        m = n // 2
        p = rng.uniform(lo, hi, size=(m, 3))
        ax = rng.integers(0, 3, size=m)
        side = rng.integers(0, 2, size=m)
        p[np.arange(m), ax] = np.where(side == 0, lo[ax], hi[ax])
        out.append(p)
    return np.concatenate(out)


def icp_apply(icp: np.ndarray, pts: np.ndarray) -> np.ndarray:
    """x_s = s*Ri*x_u + ti."""
    return pts @ icp[:3, :3].T + icp[:3, 3]


def make_reference_like_scene(name: str, sh_degree: int = 3, seed: int = 101, with_tblock: bool = False,
                              assets: Optional[Dict[str, np.ndarray]] = None) -> SplatScene:
    """robots-scene-v2 (N=113 831, 7 real xarm6-1 masks) or divar113vhw (N=292 247, 6 real masks)
    with synthetic parameters; optionally the T-block as one more link of 4 000 Gaussians
    (BASELINE.json configs 1-2)."""
    assets = assets or load_reference_assets()
    masks = unpack_masks(assets, name)
    N = masks.shape[1]
    icp = assets[f"{name}/icp"]
    surf = icp_apply(icp, assets[f"{name}/surface_points"].astype(np.float64))
    # split the surface samples into L height bands so each link's Gaussians stay together
    order = np.argsort(surf @ icp[:3, 2] / np.linalg.norm(icp[:3, 2]))
    bands = np.array_split(order, masks.shape[0])
    link_points = [surf[b] for b in bands]
    if with_tblock:
        nt = 4000
        extra = np.zeros((1, N + nt), dtype=bool)
        extra[0, N:] = True
        masks = np.concatenate([np.pad(masks, ((0, 0), (0, nt))), extra], axis=0)
        tb = tblock_points(4096, seed + 1) + np.array([0.35, 0.1, 0.0])
        link_points.append(icp_apply(icp, tb))
        N += nt
    p = synthetic_params(N, sh_degree, seed, masks, link_points)
    return canonicalize(*p, masks)


def make_synthetic_scene(N: int, sh_degree: int, seed: int, link_counts: Sequence[int] = ()) -> SplatScene:
    masks = scatter_masks(N, link_counts, seed + 17) if len(link_counts) else None
    link_points = None
    if masks is not None:
        rng = np.random.default_rng(seed + 29)
        link_points = [rng.uniform(-0.3, 0.3, size=(1, 3)) + rng.normal(0, 0.05, size=(512, 3)) for _ in link_counts]
    p = synthetic_params(N, sh_degree, seed, masks, link_points)
    return canonicalize(*p, masks)


def random_link_poses(E: int, L: int, seed: int = 1000) -> np.ndarray:
    """[E,L,12] fp32 (rot row-major | pos): rot = Exp(w), w ~ U(-0.6,0.6)^3, pos ~ U(-0.05,0.05)^3,
    composed along the chain, about each env's own seed."""
    from .transforms import SE3, SO3

    out = np.zeros((E, max(L, 1), 12), dtype=np.float32)
    for e in range(E):
        rng = np.random.default_rng(seed + e)
        T = SE3.identity()
        for l in range(L):
            step = SE3.from_rotation_and_translation(SO3.exp(rng.uniform(-0.6, 0.6, 3)), rng.uniform(-0.05, 0.05, 3))
            T = T.multiply(step) if l else step
            out[e, l, :9] = T.rotation().as_matrix().reshape(9)
            out[e, l, 9:] = T.translation()
    return out[:, :L] if L else out[:, :0]


def fov_intrinsics(W: int, H: int, vfov_deg: float = 60.0) -> np.ndarray:
    fy = H / (2.0 * np.tan(np.radians(vfov_deg) / 2.0))
    return np.array([[fy, 0.0, W / 2.0], [0.0, fy, H / 2.0], [0.0, 0.0, 1.0]], dtype=np.float64)


def lookat_viewmat(eye, target, up=(0.0, 0.0, 1.0)) -> np.ndarray:
    """World->camera 4x4, OpenCV axes (+x right, +y down, +z forward)."""
    eye, target, up = (np.asarray(v, dtype=np.float64) for v in (eye, target, up))
    z = target - eye
    z /= np.linalg.norm(z)
    x = np.cross(z, up)
    x /= np.linalg.norm(x)
    y = np.cross(z, x)
    R = np.stack([x, y, z])  # rows
    vm = np.eye(4)
    vm[:3, :3] = R
    vm[:3, 3] = -R @ eye
    return vm


def ring_cameras(n: int, target, seed: int, radius=(0.8, 1.4), elev_deg=(15.0, 60.0)) -> np.ndarray:
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        r = rng.uniform(*radius)
        el = np.radians(rng.uniform(*elev_deg))
        az = 2 * np.pi * (i + rng.uniform(0, 1)) / n
        eye = np.asarray(target) + r * np.array([np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el)])
        out.append(lookat_viewmat(eye, target))
    return np.stack(out)


def c2w_opencv_to_viewmat(wxyz, position) -> np.ndarray:
    """viser-side camera pose (camera-to-world, OpenCV axes; splat_env_wrapper.py:148-157) -> world->camera."""
    from .transforms import SE3, SO3

    T = SE3.from_rotation_and_translation(SO3(np.asarray(wxyz, dtype=np.float64)), position)
    return T.inverse().as_matrix()


def c2w_opengl_to_viewmat(c2w) -> np.ndarray:
    """nerfstudio-side pose (camera-to-world, OpenGL axes; nerfstudio_utils.py:125-138) -> world->camera
    with OpenCV axes: flip the camera's y and z columns, then invert [RECALLED splatfacto get_outputs]."""
    c = np.eye(4)
    c[:3, :4] = np.asarray(c2w, dtype=np.float64)[:3, :4]
    c[:3, 1:3] *= -1.0
    Rm, t = c[:3, :3], c[:3, 3]
    vm = np.eye(4)
    vm[:3, :3] = Rm.T
    vm[:3, 3] = -Rm.T @ t
    return vm
