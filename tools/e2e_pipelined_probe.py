"""Where do the last few per cent between `e2e` and `value` go?  Variants of bench.py's two-handle render_async loop
(run on the GPU box): how late the host consumes a step, which outputs travel, pinned-host against device poses."""
import sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench
from sim_a_splat_b200.renderer import PipelinedSplatRenderer

sc, cam, K, pose_sets = bench.build_workload(0)
I, W, H = bench.E_PER_GPU, bench.W, bench.H
pr = PipelinedSplatRenderer(sc, depth=2, device=0, max_images=I, max_envs=I, max_width=W, max_height=H,
                            max_isects=bench.CAP_ALL, max_isects_src=bench.CAP_SRC)
pr.set_cameras(np.repeat(cam, I, axis=0), K, W, H, np.arange(I))
pin = [torch.from_numpy(p).pin_memory().numpy() for p in pose_sets]
dev = [torch.from_numpy(p).cuda() for p in pose_sets]
steps = 300


def loop(names, poses, lag):
    for i in range(4):
        pr.set_link_poses(poses[i % 4]); pr.wait(pr.render_async(names))
    torch.cuda.synchronize()
    t0 = time.perf_counter(); tickets = []
    for i in range(steps):
        pr.set_link_poses(poses[i % 4])
        tickets.append(pr.render_async(names))
        if i >= lag:
            pr.wait(tickets[i - lag])
    for tk in tickets[-lag:]:
        pr.wait(tk)
    return (time.perf_counter() - t0) / steps * 1e3


def device_loop(poses):
    for i in range(4):
        pr.set_link_poses(poses[i % 4]); pr.render(("rgb_u8", "depth"))
    pr.join(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        pr.set_link_poses(poses[i % 4]); pr.render(("rgb_u8", "depth"))
    pr.join(); torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps * 1e3


for rep in range(2):
    print(f"device outputs, device poses            {device_loop(dev):.4f} ms/step", flush=True)
    print(f"device outputs, pinned poses            {device_loop(pin):.4f} ms/step", flush=True)
    for tag, args in [("host rgb_u8+depth, consume i-2 (bench)", (("rgb_u8", "depth"), pin, 2)),
                      ("host rgb_u8+depth, consume i-3", (("rgb_u8", "depth"), pin, 3)),
                      ("host rgb_u8+depth, consume i-1", (("rgb_u8", "depth"), pin, 1)),
                      ("host rgb_u8+depth_f16, consume i-2", (("rgb_u8", "depth_f16"), pin, 2)),
                      ("host rgb_u8 only, consume i-2", (("rgb_u8",), pin, 2)),
                      ("host rgb_u8+depth, device poses, i-2", (("rgb_u8", "depth"), dev, 2))]:
        print(f"{tag:40s}{loop(*args):.4f} ms/step", flush=True)
pr.close()
