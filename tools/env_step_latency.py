#!/usr/bin/env python
"""Wall-clock latency of the reference-shaped single-env path: SplatEnvWrapper.step() with the camera setup the
reference's scripts use (2 cameras, 240x320: /root/reference/tests/test_splat_wrapper.py:77-101), on the toy
assets of tests/compat_fixtures.py scaled to the robots-scene-v2 Gaussian count.  Prints one JSON line."""
import json
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from compat_fixtures import FakeManipulatorEnv, camera_setup, make_assets  # noqa: E402
from sim_a_splat_b200.env.splat.splat_env_wrapper import SplatEnvWrapper  # noqa: E402
from sim_a_splat_b200.scene import load_reference_assets  # noqa: E402


def main():
    with tempfile.TemporaryDirectory() as d:
        info = make_assets(Path(d), load_reference_assets(), n_gaussians=113_831, sh_degree=3)
        env = SplatEnvWrapper(FakeManipulatorEnv(info, seed=1), info["root"], info["obj"], info["cfg"],
                              task_assets_path=info["task_assets_path"], task_assets_name=info["task_assets_name"])
        env._configure_cameras(camera_setup((240, 320)))
        env.reset(seed=2)
        a = np.array([0.01, -0.01, 0.001])
        for _ in range(10):
            env.step(a)
        ts = []
        for _ in range(100):
            t0 = time.perf_counter()
            obs, *_ = env.step(a)
            ts.append(time.perf_counter() - t0)
        t_noobs = []
        for _ in range(100):
            t0 = time.perf_counter()
            env.step(a, noobs=True)
            t_noobs.append(time.perf_counter() - t0)
        print(json.dumps({"what": "SplatEnvWrapper.step, 1 env, 2 cameras 240x320 (1 moving + 1 fixed), N=113831+4000 SH3",
                          "step_ms_median": round(float(np.median(ts)) * 1e3, 3), "step_ms_p90": round(float(np.percentile(ts, 90)) * 1e3, 3),
                          "step_noobs_ms_median": round(float(np.median(t_noobs)) * 1e3, 3),
                          "obs_keys": sorted(obs.keys()), "camera_shape": list(obs["camera_0"].shape)}))
        env.close()


if __name__ == "__main__":
    main()
