#!/bin/bash
# sweep of SAS_COMPACT_MAX (per-tile gate of the compacted compositor lists) on the headline batch and the configs[3] shard
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout=900 -x -k "configs or parity or api_r2" > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?: $(tail -1 gpurun_out/pytest_gpu.log)"
for m in 32 28 24 20 12; do
  SAS_COMPACT_MAX=$m timeout 300 python bench.py --steps 100 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/bench_m$m.log 2> gpurun_out/bench_m$m.err
  python - gpurun_out/bench_m$m.log $m <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("max", sys.argv[2], round(d["value"]), "fps", round(d["ms_per_step"],4), "serial", round(d["serial"]["ms_per_step"],4), {k:v["ms"] for k,v in d["stages"].items()}, "staged", d["fp32"]["staged_gaussians_per_step"])
PY
  SAS_COMPACT_MAX=$m timeout 600 python tools/bench_configs.py --steps 10 --cams5 8 --only 3 > gpurun_out/configs_m$m.jsonl 2> gpurun_out/configs_m$m.err
  python - gpurun_out/configs_m$m.jsonl $m <<'PY'
import json,sys
for l in open(sys.argv[1]):
    d=json.loads(l); print("   max %s"%sys.argv[2], d["config"][:40], d["ms_per_step"], d["stages_ms"])
PY
done
