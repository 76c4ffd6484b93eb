"""Compositor stage time against the number of envs of the pusht batch: a constant floor = latency of the heaviest
tile's chain (tail-bound), a slope = throughput-bound.  GPU."""
import sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench
from sim_a_splat_b200 import scene as S
from sim_a_splat_b200.renderer import SplatRenderer

sc, cam, K, _ = bench.build_workload(0)
W, H = bench.W, bench.H
for E in [int(a) for a in sys.argv[1:]] or [16, 32, 64, 128, 256, 512, 1024]:
    r = SplatRenderer(sc, device=0, max_images=E, max_envs=E, max_width=W, max_height=H, max_isects=48_000_000 * max(1, E // 256))
    r.set_cameras(np.repeat(cam, E, axis=0), K, W, H, np.arange(E))
    poses = [torch.from_numpy(S.random_link_poses(E, sc.L, seed=1000 + 100_000 * s)).cuda() for s in range(2)]
    out = {n: r._alloc(n, False) for n in ("rgb", "depth", "alpha")}
    for i in range(3):
        r.set_link_poses(poses[i % 2]); r.render(("rgb", "depth", "alpha"), out=out, check=False)
    assert r.stats()["overflow"] == 0
    r.set_stage_timing(True)
    acc = {k: 0.0 for k in r.STAGES}
    n = 20
    for i in range(n):
        r.set_link_poses(poses[i % 2]); r.render(("rgb", "depth", "alpha"), out=out, check=False)
        for k, v in r.stage_ms().items():
            acc[k] += v / n
    st = r.stats()
    print(f"E={E:5d} composite {acc['composite']:.3f} sort {acc['sort']:.3f} project {acc['project']:.3f} emit {acc['emit']:.3f} scan {acc['scan']:.3f} "
          f"| total {sum(acc.values()):.3f} ms | staged {st['n_staged']} max_list {st['max_tile_list']}", flush=True)
    r.close()
