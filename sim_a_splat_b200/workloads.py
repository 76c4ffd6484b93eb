"""The five BASELINE.json configs as seeded synthetic workloads: real Gaussian counts, real
link-mask structure, real ICP / camera metadata of the reference's assets (tests/golden/assets.npz), synthetic
Gaussian parameters -- the trained checkpoints are git-LFS stubs.  Used by bench.py (configs[1] is the bench line,
the others its `other_configs` block and the strong-scaling entry), tools/bench_configs.py and the GPU tests.

Every function returns (scene, viewmats [I,4,4], K [3,3], W, H, link_poses [E,L,12] or None, env_of_image [I] or None).
"""
from __future__ import annotations

import numpy as np

from . import scene as S


def config0_single_frame(assets=None):
    """configs[0]: the pusht scene (robots-scene-v2 + T-block), one 640x480 frame -- the CPU-runnable case."""
    sc = S.make_reference_like_scene("robots-scene-v2", 3, 101, with_tblock=True, assets=assets)
    cen = sc.means[sc.link_id >= 0].mean(0)
    return sc, S.ring_cameras(1, cen, 7), S.fov_intrinsics(640, 480), 640, 480, S.random_link_poses(1, sc.L, 1000), np.zeros(1, np.int32)


def config1_pusht(n_envs: int = 256, pose_seed: int = 1000, assets=None):
    """configs[1]: pusht gym env, n_envs x 96x96, one fixed camera, per-env masked link poses (the bench line)."""
    sc = S.make_reference_like_scene("robots-scene-v2", 3, 102, with_tblock=True, assets=assets)
    cen = sc.means[sc.link_id >= 0].mean(0)
    cam = S.ring_cameras(1, cen, 8)
    return (sc, np.repeat(cam, n_envs, axis=0), S.fov_intrinsics(96, 96), 96, 96, S.random_link_poses(n_envs, sc.L, seed=pose_seed),
            np.arange(n_envs, dtype=np.int32))


def config2_xarm(assets=None):
    """configs[2]: robots-scene-v2 xarm, real xarm6-1 masks (L = 7), 4 cameras x 1280x720, one env."""
    sc = S.make_reference_like_scene("robots-scene-v2", 3, 103, assets=assets)
    cen = sc.means[sc.link_id >= 0].mean(0)
    return sc, S.ring_cameras(4, cen, 9), S.fov_intrinsics(1280, 720), 1280, 720, S.random_link_poses(1, 7, 2000), np.zeros(4, np.int32)


def config3_divar(n_envs: int = 2048, env_lo: int = 0, env_hi: int = None, assets=None):
    """configs[3]: divar113vhw (N = 298 588 canonical, L = 6), n_envs x 2 fixed cameras x 224x224.  env_lo / env_hi
    select a rank's shard of the n_envs (poses are drawn for all n_envs so that a shard is a slice of the full job)."""
    sc = S.make_reference_like_scene("divar113vhw", 3, 104, assets=assets)
    cen = sc.means[sc.link_id >= 0].mean(0)
    env_hi = n_envs if env_hi is None else env_hi
    E = env_hi - env_lo
    cams = S.ring_cameras(2, cen, 10)
    poses = S.random_link_poses(n_envs, 6, 3000)[env_lo:env_hi]
    return (sc, np.tile(cams, (E, 1, 1)), S.fov_intrinsics(224, 224), 224, 224, np.ascontiguousarray(poses),
            np.repeat(np.arange(E, dtype=np.int32), 2))


def config4_stress(n_cameras: int = 64, n_gaussians: int = 6_000_000, first_camera: int = 0):
    """configs[4]: synthetic 6 M-Gaussian SH-3 scene at 1920x1080, cameras [first_camera, first_camera + n_cameras) of
    the 64-camera ring (sort / composite stress; no links)."""
    sc = S.make_synthetic_scene(n_gaussians, 3, seed=105)
    cams = S.ring_cameras(64, (0, 0, 0), seed=14)[first_camera:first_camera + n_cameras]
    return sc, cams, S.fov_intrinsics(1920, 1080), 1920, 1080, None, None
