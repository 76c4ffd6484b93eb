#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?: $(tail -1 gpurun_out/pytest_gpu.log)"
[ "$(tail -1 gpurun_out/pytest_gpu.log | grep -c failed)" != "0" ] && tail -60 gpurun_out/pytest_gpu.log
timeout 900 python tools/fuzz_parity.py 60 17000 > gpurun_out/fuzz.log 2>&1
echo "fuzz exit $?: $(tail -1 gpurun_out/fuzz.log)"
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], round(d["value"]), "fps", round(d["ms_per_step"],3), "ms/step; serial", round(d["serial"]["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), {k:v["ms"] for k,v in d["stages"].items()}, "launches", d["launches_per_step"])
except Exception as e:
    print(sys.argv[2], "no bench line", e)
PY
}
for v in 1 0; do
  SAS_SORT_BIG=$v timeout 300 python bench.py --steps 100 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/bench_b$v.log 2> gpurun_out/bench_b$v.err; show gpurun_out/bench_b$v.log SAS_SORT_BIG=$v
done
for v in 1 0; do
  SAS_SORT_BIG=$v timeout 600 python tools/bench_configs.py --steps 10 --cams5 8 > gpurun_out/configs_b$v.jsonl 2> gpurun_out/configs_b$v.err
  python - gpurun_out/configs_b$v.jsonl $v <<'PY'
import json,sys
for l in open(sys.argv[1]):
    d=json.loads(l); print("   BIG=%s"%sys.argv[2], d["config"][:40], d["ms_per_step"], d["stages_ms"])
PY
done
# the regime the path is for: a small image over a big scene (every tile a giant bucket)
for v in 1 0; do
SAS_SORT_BIG=$v python - <<'PY'
import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
from sim_a_splat_b200 import scene as S
from sim_a_splat_b200.renderer import SplatRenderer
sc = S.make_synthetic_scene(2_000_000, 3, seed=105)
W = H = 96
r = SplatRenderer(sc, max_images=2, max_width=W, max_height=H)
r.set_cameras(S.ring_cameras(2, (0, 0, 0), seed=14), S.fov_intrinsics(W, H), W, H)
out = {n: r._alloc(n, False) for n in ("rgb", "depth", "alpha")}
r.render(("rgb", "depth", "alpha"), out=out, check=True)
for _ in range(3): r.render(("rgb", "depth", "alpha"), out=out, check=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): r.render(("rgb", "depth", "alpha"), out=out, check=False)
e1.record(); torch.cuda.synchronize()
r.set_stage_timing(True); r.render(("rgb", "depth", "alpha"), out=out, check=False)
st = r.stats()
print("small image over a big scene: SAS_SORT_BIG=%s  %.3f ms/step" % (os.environ["SAS_SORT_BIG"], e0.elapsed_time(e1) / 10), {k: round(v, 3) for k, v in r.stage_ms().items()}, "max list", st["max_tile_list"], "M", st["n_isects"])
PY
done
