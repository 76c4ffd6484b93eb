#!/bin/bash
mkdir -p gpurun_out
for ratio in 1 4; do
  SAS_COMPACT_RATIO=$ratio timeout 600 python tools/bench_configs.py --steps 20 --cams5 8 --only 0,2,3 > gpurun_out/configs_r$ratio.jsonl 2> gpurun_out/configs_r$ratio.err
  python - gpurun_out/configs_r$ratio.jsonl $ratio <<'PY'
import json,sys
for l in open(sys.argv[1]):
    d=json.loads(l); print("   ratio %s"%sys.argv[2], d["config"][:40], d["ms_per_step"], d["stages_ms"])
PY
done
timeout 300 python bench.py --steps 100 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/bench.log 2> gpurun_out/bench.err
python - gpurun_out/bench.log <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(round(d["value"]), "fps", round(d["ms_per_step"],4), "serial", round(d["serial"]["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), {k:v["ms"] for k,v in d["stages"].items()}, "staged", d["fp32"]["staged_gaussians_per_step"])
PY
SAS_COMPACT_RATIO=1 timeout 900 python -m pytest tests -m gpu -q --timeout=900 -x -k "parity or compat or configs" > gpurun_out/pytest_gpu.log 2>&1
echo "pytest (ratio 1) exit $?: $(tail -1 gpurun_out/pytest_gpu.log)"
