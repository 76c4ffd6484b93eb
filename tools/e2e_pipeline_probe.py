"""Where does the pipelined e2e step time go?  Variants of bench.py's render_async loop (run on the GPU box)."""
import sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench
from sim_a_splat_b200.renderer import SplatRenderer

sc, cam, K, pose_sets = bench.build_workload(0)
I, W, H = bench.E_PER_GPU, bench.W, bench.H
r = SplatRenderer(sc, device=0, max_images=I, max_envs=I, max_width=W, max_height=H, max_isects=48_000_000)
r.set_cameras(np.repeat(cam, I, axis=0), K, W, H, np.arange(I))
pin = [torch.from_numpy(p).pin_memory() for p in pose_sets]
dev = [torch.from_numpy(p).cuda() for p in pose_sets]
steps = 200

def loop(names, poses, use_stream, depth2=True):
    s = torch.cuda.Stream() if use_stream else None
    ctx = torch.cuda.stream(s) if use_stream else torch.cuda.stream(torch.cuda.current_stream())
    with ctx:
        for i in range(3):
            r.set_link_poses(poses[i]); r.wait(r.render_async(names))
        torch.cuda.synchronize()
        t0 = time.perf_counter(); prev = None
        for i in range(steps):
            r.set_link_poses(poses[i % 4])
            tk = r.render_async(names)
            if depth2:
                if prev is not None: r.wait(prev)
            else:
                r.wait(tk)
            prev = tk
        r.wait(prev)
        dt = time.perf_counter() - t0
    return dt / steps * 1e3

pn = [p.numpy() for p in pin]
for tag, args in [("bench loop (default stream, pinned poses)", (("rgb_u8", "depth"), pn, False)),
                  ("explicit stream", (("rgb_u8", "depth"), pn, True)),
                  ("device poses", (("rgb_u8", "depth"), dev, False)),
                  ("rgb_u8 only", (("rgb_u8",), pn, False)),
                  ("depth 1 (wait at once)", (("rgb_u8", "depth"), pn, False, False))]:
    print(f"{tag:45s} {loop(*args):.3f} ms/step", flush=True)
# host-side cost of issuing one step (no wait)
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(50):
    r.set_link_poses(pn[i % 4]); r.render_async(("rgb_u8", "depth"))
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"issue cost {1e3*(t1-t0)/50:.3f} ms/step, drained after {1e3*(t2-t0)/50:.3f} ms/step")
r.close()
