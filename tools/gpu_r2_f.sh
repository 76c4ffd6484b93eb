#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?: $(tail -1 gpurun_out/pytest_gpu.log)"
timeout 600 python tools/fuzz_parity.py 40 14000 > gpurun_out/fuzz.log 2>&1
echo "fuzz exit $?: $(tail -1 gpurun_out/fuzz.log)"
for v in 2 3 4; do
  SAS_PROJECT_D=$v timeout 600 python tools/bench_configs.py --steps 10 --cams5 8 > gpurun_out/configs_d$v.jsonl 2> gpurun_out/configs_d$v.err
  echo "configs SAS_PROJECT_D=$v exit $?"; python - gpurun_out/configs_d$v.jsonl <<'PY'
import json,sys
for l in open(sys.argv[1]):
    d=json.loads(l); print(d["config"][:40], d["ms_per_step"], d["stages_ms"])
PY
done
timeout 600 python bench.py --steps 100 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/bench.log 2> gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print("bench", round(d["value"]), "serial", round(d["serial"]["value"]), "e2e", round(d["e2e"]["value"]), {k:v["ms"] for k,v in d["stages"].items()})
PY
