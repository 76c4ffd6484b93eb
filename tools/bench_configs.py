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
    from sim_a_splat_b200 import workloads as WL

    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--n5", type=int, default=6_000_000)
    ap.add_argument("--cams5", type=int, default=8)
    ap.add_argument("--only", type=str, default="")
    a = ap.parse_args()
    only = [int(x) for x in a.only.split(",")] if a.only else [0, 2, 3, 4]
    if 0 in only:
        sc, vms, K, W, H, poses, env = WL.config0_single_frame()
        time_config("configs[0] tblock pusht scene 640x480 x1", sc, vms, K, W, H, poses, env, a.steps)
    if 2 in only:
        sc, vms, K, W, H, poses, env = WL.config2_xarm()
        time_config("configs[2] xarm 4 cameras x 1280x720", sc, vms, K, W, H, poses, env, a.steps)
    if 3 in only:  # the 256-env x 2-camera shard one of 8 GPUs renders
        sc, vms, K, W, H, poses, env = WL.config3_divar(2048, 0, 256)
        time_config("configs[3] divar113vhw shard: 256 envs x 2 cameras x 224x224", sc, vms, K, W, H, poses, env, a.steps, max_isects=120_000_000)
    if 4 in only:
        sc, vms, K, W, H, poses, env = WL.config4_stress(a.cams5, a.n5)
        time_config(f"configs[4] synthetic {a.n5} Gaussians SH3 1920x1080 x{a.cams5} cameras (one chunk of the 64)", sc, vms, K, W, H,
                    None, None, max(a.steps // 4, 3))


if __name__ == "__main__":
    main()
