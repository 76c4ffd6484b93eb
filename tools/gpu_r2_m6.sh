#!/bin/bash
# Morton order on by default for scenes >= 2 M background Gaussians: the 6 M-Gaussian tests + the configs[4] chunk, A/B
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_configs_gpu.py -m gpu -q --timeout=900 -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?: $(tail -1 gpurun_out/pytest_gpu.log)"
for m in auto 0; do
  if [ $m = auto ]; then unset SAS_MORTON; else export SAS_MORTON=$m; fi
  timeout 900 python tools/bench_configs.py --steps 5 --cams5 8 --only 4 > gpurun_out/configs_m$m.jsonl 2> gpurun_out/configs_m$m.err
  python - gpurun_out/configs_m$m.jsonl $m <<'PY'
import json,sys
for l in open(sys.argv[1]):
    d=json.loads(l); print("   SAS_MORTON=%s"%sys.argv[2], d["config"][:40], d["ms_per_step"], d["stages_ms"])
PY
done
