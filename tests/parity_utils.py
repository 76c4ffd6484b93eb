"""Shared helpers of the GPU parity tests: run the CUDA path (through the C ABI) and the CPU
oracle on the same seeded inputs and compare every intermediate."""
from __future__ import annotations

import numpy as np

import oracle
from sim_a_splat_b200 import scene as S
from sim_a_splat_b200.renderer import SplatRenderer

INT_KEYS = ("radii", "tile_rect", "tiles_per_gauss", "isect_offsets", "isect_ids", "flatten_ids")
FLOAT_KEYS = ("means2d", "depths", "conics", "colors")
IMG_KEYS = ("rgb", "depth", "alpha")


def psnr(a, b):
    mse = float(np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2))
    return np.inf if mse == 0 else 10.0 * np.log10(1.0 / mse)


def run_both(scene: S.SplatScene, viewmats, Ks, W, H, poses=None, env_of_image=None, bg=(0.0, 0.0, 0.0), renderer=None,
             max_isects=0):
    viewmats = np.asarray(viewmats, dtype=np.float32).reshape(-1, 4, 4)
    I = viewmats.shape[0]
    Ks = np.asarray(Ks, dtype=np.float32).reshape(-1, 3, 3)
    if Ks.shape[0] == 1 and I > 1:
        Ks = np.repeat(Ks, I, axis=0)
    E = 1 if poses is None else np.asarray(poses).reshape(-1, max(scene.L, 1), 12).shape[0]
    own = renderer is None
    r = renderer or SplatRenderer(scene, max_images=I, max_envs=E, max_width=W, max_height=H, max_isects=max_isects)
    if scene.L:
        r.set_link_poses(poses)
    r.set_cameras(viewmats, Ks, W, H, env_of_image)
    out = r.render(outputs=("rgb", "depth", "alpha", "rgb_u8", "n_contrib"), background=bg)
    gpu = {k: v.cpu().numpy() for k, v in out.items()}
    gpu.update(r.debug_intermediates())
    if own:
        r.close()
    ref = oracle.render(scene.as_oracle_dict(), viewmats, Ks, W, H, poses, env_of_image, bg=bg)
    return gpu, ref


def compare(gpu, ref) -> dict:
    """Mismatch report: exact-equality counts for the integer stages, max-abs / bit-exactness for floats."""
    rep = {"M_gpu": int(gpu["M"]), "M_ref": int(ref["M"])}
    for k in INT_KEYS:
        a, b = np.asarray(gpu[k]), np.asarray(ref[k])
        rep[k] = "shape %s vs %s" % (a.shape, b.shape) if a.shape != b.shape else int(np.sum(a != b))
    for k in FLOAT_KEYS:
        a, b = np.asarray(gpu[k]), np.asarray(ref[k])
        rep[k + "_maxabs"] = float(np.max(np.abs(a - b))) if a.size else 0.0
        rep[k + "_bitdiff"] = int(np.sum(a.view(np.uint32) != b.view(np.uint32)))
        rep[k + "_scale"] = float(np.max(np.abs(b))) if b.size else 0.0
    rep["n_contrib"] = int(np.sum(gpu["n_contrib"] != ref["n_contrib"]))
    for k in IMG_KEYS:
        rep[k + "_maxabs"] = float(np.max(np.abs(gpu[k] - ref[k])))
        rep[k + "_bitdiff"] = int(np.sum(gpu[k].view(np.uint32) != ref[k].view(np.uint32)))
    rep["psnr_rgb"] = psnr(gpu["rgb"], ref["rgb"])
    u8 = np.floor(np.clip(ref["rgb"], 0.0, 1.0) * np.float32(255.0) + np.float32(0.5)).astype(np.uint8)
    rep["rgb_u8"] = int(np.sum(gpu["rgb_u8"] != u8))
    return rep


def assert_parity(rep: dict, img_tol: float = 1e-3):
    """The bar of BASELINE.json north_star: tile ids, sort order and per-pixel contributor counts
    bit-exact; RGB / depth within max-abs 1e-3 (fp32), PSNR > 60 dB."""
    assert rep["M_gpu"] == rep["M_ref"], rep
    for k in INT_KEYS:
        assert rep[k] == 0, (k, rep)
    assert rep["n_contrib"] == 0, rep
    for k in FLOAT_KEYS:
        assert rep[k + "_maxabs"] <= 1e-5 * max(1.0, rep[k + "_scale"]), (k, rep)
    for k in IMG_KEYS:
        assert rep[k + "_maxabs"] <= img_tol, (k, rep)
    assert rep["psnr_rgb"] > 60.0, rep
    assert rep["rgb_u8"] == 0, rep
