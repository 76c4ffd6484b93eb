#!/usr/bin/env python
"""Step times of the BASELINE.json configs that are NOT the bench line (configs[0], [2], [3] shard, [4]): one
JSON object per config, for the table in DESIGN.md.  bench.py stays the contract for configs[1].

    python tools/bench_configs.py [--steps 20] [--n5 6000000] [--cams5 8]
"""
import argparse
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from sim_a_splat_b200 import scene as S  # noqa: E402
from sim_a_splat_b200.renderer import SplatRenderer  # noqa: E402


def time_config(name, sc, vms, K, W, H, poses, env, steps, max_isects=0):
    I = vms.shape[0]
    E = 1 if poses is None else poses.shape[0]
    r = SplatRenderer(sc, max_images=I, max_envs=E, max_width=W, max_height=H, max_isects=max_isects)
    if sc.L:
        r.set_link_poses(torch.from_numpy(poses).cuda())
    r.set_cameras(vms, K, W, H, env)
    names = ("rgb", "depth", "alpha")
    out = {n: r._alloc(n, False) for n in names}
    r.render(names, out=out, check=True)
    pd = torch.from_numpy(poses).cuda() if sc.L else None
    for _ in range(3):
        r.render(names, out=out, check=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        if sc.L:
            r.set_link_poses(pd)
        r.render(names, out=out, check=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    r.set_stage_timing(True)
    r.render(names, out=out, check=False)
    stage = r.stage_ms()
    st = r.stats()
    r.close()
    line = {"config": name, "images": I, "width": W, "height": H, "n_gaussians": sc.N, "ms_per_step": round(ms, 4),
            "frames_per_s": round(I / ms * 1e3, 1), "stages_ms": {k: round(v, 4) for k, v in stage.items()},
            "n_views": st["n_views"], "n_isects": st["n_isects"], "n_isects_src": st["n_isects_src"], "max_tile_list": st["max_tile_list"]}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--n5", type=int, default=6_000_000)
    ap.add_argument("--cams5", type=int, default=8)
    a = ap.parse_args()
    # configs[0]: one 640x480 frame of the pusht scene (the reference's CPU-runnable case)
    sc = S.make_reference_like_scene("robots-scene-v2", 3, 101, with_tblock=True)
    cen = sc.means[sc.link_id >= 0].mean(0)
    time_config("configs[0] tblock pusht scene 640x480 x1", sc, S.ring_cameras(1, cen, 7), S.fov_intrinsics(640, 480), 640, 480,
                S.random_link_poses(1, sc.L, 1000), np.zeros(1, np.int32), a.steps)
    # configs[2]: xarm, 4 cameras x 1280x720
    sc = S.make_reference_like_scene("robots-scene-v2", 3, 103)
    cen = sc.means[sc.link_id >= 0].mean(0)
    time_config("configs[2] xarm 4 cameras x 1280x720", sc, S.ring_cameras(4, cen, 9), S.fov_intrinsics(1280, 720), 1280, 720,
                S.random_link_poses(1, 7, 2000), np.zeros(4, np.int32), a.steps)
    # configs[3]: the 256-env x 2-camera shard one of 8 GPUs renders
    sc = S.make_reference_like_scene("divar113vhw", 3, 104)
    cen = sc.means[sc.link_id >= 0].mean(0)
    E = 256
    time_config("configs[3] divar113vhw shard: 256 envs x 2 cameras x 224x224", sc, np.tile(S.ring_cameras(2, cen, 10), (E, 1, 1)),
                S.fov_intrinsics(224, 224), 224, 224, S.random_link_poses(E, 6, 3000), np.repeat(np.arange(E, dtype=np.int32), 2), a.steps,
                max_isects=120_000_000)
    # configs[4]: synthetic N-Gaussian SH-3 scene at 1920x1080, cameras in chunks
    sc = S.make_synthetic_scene(a.n5, 3, seed=105)
    time_config(f"configs[4] synthetic {a.n5} Gaussians SH3 1920x1080 x{a.cams5} cameras (one chunk of the 64)", sc,
                S.ring_cameras(a.cams5, (0, 0, 0), seed=14), S.fov_intrinsics(1920, 1080), 1920, 1080, None, None, max(a.steps // 4, 3))


if __name__ == "__main__":
    main()
