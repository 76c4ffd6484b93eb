"""Does overlapping two independent renders (env groups on two streams) hide the compositor's tail?  GPU."""
import sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench
from sim_a_splat_b200 import scene as S
from sim_a_splat_b200.renderer import SplatRenderer

sc, cam, K, _ = bench.build_workload(0)
W, H = bench.W, bench.H
names = ("rgb", "depth", "alpha")

def make(E, seed):
    r = SplatRenderer(sc, device=0, max_images=E, max_envs=E, max_width=W, max_height=H, max_isects=48_000_000)
    r.set_cameras(np.repeat(cam, E, axis=0), K, W, H, np.arange(E))
    poses = [torch.from_numpy(S.random_link_poses(E, sc.L, seed=seed + 100_000 * s)).cuda() for s in range(4)]
    out = {n: r._alloc(n, False) for n in names}
    return r, poses, out

def run(groups, steps=100):
    streams = [torch.cuda.Stream() for _ in groups]
    def step(i):
        for (r, poses, out), s in zip(groups, streams):
            r.set_link_poses(poses[i % 4], stream=s)
            r.render(names, out=out, check=False, stream=s)
    for i in range(4):
        step(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        step(i)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps * 1e3

one = [make(256, 1000)]
print(f"1 x 256 envs, one stream:            {run(one):.3f} ms per 256 envs", flush=True)
two = [make(128, 1000), make(128, 5000)]
print(f"2 x 128 envs, two streams:           {run(two):.3f} ms per 256 envs", flush=True)
four = two + [make(128, 7000), make(128, 9000)]
print(f"4 x 128 envs, four streams:          {run(four) / 2:.3f} ms per 256 envs", flush=True)
twob = [one[0], make(256, 3000)]
print(f"2 x 256 envs, two streams:           {run(twob) / 2:.3f} ms per 256 envs", flush=True)
