#!/usr/bin/env python
"""configs[3] shard (divar113vhw, 256 envs x 2 cameras x 224x224) for an ncu launch list: 3 renders."""
import sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from sim_a_splat_b200 import scene as S
from sim_a_splat_b200.renderer import SplatRenderer
sc = S.make_reference_like_scene("divar113vhw", 3, 104)
cen = sc.means[sc.link_id >= 0].mean(0)
E = 256
r = SplatRenderer(sc, max_images=2 * E, max_envs=E, max_width=224, max_height=224, max_isects=120_000_000)
r.set_link_poses(S.random_link_poses(E, 6, 3000))
r.set_cameras(np.tile(S.ring_cameras(2, cen, 10), (E, 1, 1)), S.fov_intrinsics(224, 224), 224, 224, np.repeat(np.arange(E, dtype=np.int32), 2))
out = {n: r._alloc(n, False) for n in ("rgb", "depth", "alpha")}
for _ in range(3):
    r.render(("rgb", "depth", "alpha"), out=out, check=False)
torch.cuda.synchronize()
print(r.stats())
