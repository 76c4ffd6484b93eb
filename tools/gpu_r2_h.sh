#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?: $(tail -1 gpurun_out/pytest_gpu.log)"
[ "$(tail -1 gpurun_out/pytest_gpu.log | grep -c failed)" != "0" ] && tail -40 gpurun_out/pytest_gpu.log
timeout 600 python tools/fuzz_parity.py 40 15000 > gpurun_out/fuzz.log 2>&1
echo "fuzz exit $?: $(tail -1 gpurun_out/fuzz.log)"
timeout 600 python tools/bench_configs.py --steps 10 --cams5 8 > gpurun_out/configs.jsonl 2> gpurun_out/configs.err
python - gpurun_out/configs.jsonl <<'PY'
import json,sys
for l in open(sys.argv[1]):
    d=json.loads(l); print(d["config"][:40], d["ms_per_step"], d["stages_ms"])
PY
bash tools/gpu_r2_g.sh
