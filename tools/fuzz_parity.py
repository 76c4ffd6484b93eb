#!/usr/bin/env python
"""Randomised parity sweep: random scene sizes, SH degrees, link structures, image sizes (ragged), camera
placements (inside the scene, grazing, far), env/camera sharing patterns, Gaussian scale mixes (incl. screen-filling
ones) -- every intermediate and every image against the CPU oracle.  Prints one line per case; exit code 1 on the
first mismatch.   python tools/fuzz_parity.py [n_cases] [seed0]"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from parity_utils import assert_parity, compare, run_both  # noqa: E402
from sim_a_splat_b200 import scene as S  # noqa: E402


def one_case(seed: int):
    rng = np.random.default_rng(seed)
    N = int(rng.choice([300, 2000, 9000, 30000]))
    deg = int(rng.integers(0, 4))
    L = int(rng.choice([0, 1, 3, 6]))
    counts = tuple(int(c) for c in rng.integers(max(N // 60, 5), max(N // 12, 6), size=L))
    sc = S.make_synthetic_scene(N, deg, seed=seed, link_counts=counts)
    # scale mix: a few huge (screen-filling) and a few tiny Gaussians
    k = max(N // 200, 2)
    sc.scales[rng.choice(N, k, replace=False)] *= np.float32(rng.uniform(8, 40))
    sc.scales[rng.choice(N, k, replace=False)] *= np.float32(0.05)
    sc.opacities[rng.choice(N, k, replace=False)] = np.float32(0.004)   # below 1/255 after any falloff
    W = int(rng.choice([17, 64, 96, 130, 224, 333, 640]))
    H = int(rng.choice([16, 48, 96, 97, 200, 360]))
    E = int(rng.choice([1, 2, 5])) if L else 1
    n_fixed = int(rng.choice([0, 1, 2]))
    n_moving = int(rng.choice([0, 1])) if n_fixed else 1
    fixed = S.ring_cameras(max(n_fixed, 1), (0, 0, 0), seed=seed + 1, radius=(0.2, 2.5), elev_deg=(-30.0, 80.0))
    moving = S.ring_cameras(E, (0, 0, 0), seed=seed + 2, radius=(0.05, 1.5), elev_deg=(-60.0, 60.0))
    vms, env = [], []
    for e in range(E):
        for c in range(n_fixed):
            vms.append(fixed[c]); env.append(e)
        for c in range(n_moving):
            vms.append(moving[e]); env.append(e)
    vms = np.stack(vms)
    poses = S.random_link_poses(E, L, seed=seed + 3) if L else None
    fov = float(rng.choice([35.0, 60.0, 100.0]))
    bg = tuple(float(v) for v in rng.uniform(0, 1, 3))
    gpu, ref = run_both(sc, vms, S.fov_intrinsics(W, H, fov)[None], W, H, poses, env if L else None, bg=bg)
    rep = compare(gpu, ref)
    assert_parity(rep)
    return dict(seed=seed, N=N, deg=deg, L=L, E=E, I=len(vms), W=W, H=H, fov=fov, M=rep["M_ref"], views=gpu["stats"]["n_views"],
                max_list=gpu["stats"]["max_tile_list"], img_bitdiff=rep["rgb_bitdiff"] + rep["depth_bitdiff"] + rep["alpha_bitdiff"])


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    s0 = int(sys.argv[2]) if len(sys.argv) > 2 else 9000
    for k in range(n):
        try:
            print(one_case(s0 + k), flush=True)
        except AssertionError as e:
            print("MISMATCH at seed", s0 + k, str(e)[:2000], flush=True)
            sys.exit(1)
    print(f"fuzz ok: {n} cases")


if __name__ == "__main__":
    main()
