#!/bin/bash
# does compaction pay when every image has its own view (configs[0], configs[2])?  SAS_COMPACT_RATIO = 1 vs 4
mkdir -p gpurun_out
for ratio in 1 4 1 4; do
  SAS_COMPACT_RATIO=$ratio timeout 600 python tools/bench_configs.py --steps 20 --cams5 8 --only 0,2 > gpurun_out/configs_r$ratio.jsonl 2> gpurun_out/configs_r$ratio.err
  python - gpurun_out/configs_r$ratio.jsonl $ratio <<'PY'
import json,sys
for l in open(sys.argv[1]):
    d=json.loads(l); print("   ratio %s"%sys.argv[2], d["config"][:40], d["ms_per_step"], d["stages_ms"])
PY
done
SAS_COMPACT_RATIO=1 timeout 900 python -m pytest tests -m gpu -q --timeout=900 -x -k "parity or compat" > gpurun_out/pytest_gpu.log 2>&1
echo "pytest (ratio 1) exit $?: $(tail -1 gpurun_out/pytest_gpu.log)"
