"""Device-resident throughput of PipelinedSplatRenderer against its depth (bench.py's `value` loop).  GPU."""
import sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench
from sim_a_splat_b200.renderer import PipelinedSplatRenderer

sc, cam, K, pose_sets = bench.build_workload(0)
I, W, H = bench.E_PER_GPU, bench.W, bench.H
poses = [torch.from_numpy(p).cuda() for p in pose_sets]
for depth in [int(a) for a in sys.argv[1:]] or [1, 2, 3, 4]:
    pr = PipelinedSplatRenderer(sc, depth=depth, device=0, max_images=I, max_envs=I, max_width=W, max_height=H, max_isects=48_000_000)
    pr.set_cameras(np.repeat(cam, I, axis=0), K, W, H, np.arange(I))
    def step(i):
        pr.set_link_poses(poses[i % 4]); pr.render(("rgb", "depth", "alpha"))
    for i in range(8):
        step(i)
    pr.join(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(200):
        step(i)
    pr.join(); e1.record(); torch.cuda.synchronize()
    print(f"depth {depth}: {e0.elapsed_time(e1) / 200:.3f} ms/step", flush=True)
    pr.close()
