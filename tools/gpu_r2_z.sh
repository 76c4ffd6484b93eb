#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=900 -x -k "parity or compat or configs or api_r2" > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?: $(tail -1 gpurun_out/pytest_gpu.log)"
timeout 900 python tools/fuzz_parity.py 40 36000 > gpurun_out/fuzz.log 2>&1
echo "fuzz exit $?: $(tail -1 gpurun_out/fuzz.log)"
for k in 1 2; do
timeout 300 python bench.py --steps 100 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/bench_$k.log 2> gpurun_out/bench_$k.err
python - gpurun_out/bench_$k.log <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(round(d["value"]), "fps", round(d["ms_per_step"],4), "serial", round(d["serial"]["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), {k:v["ms"] for k,v in d["stages"].items()}, "staged", d["fp32"]["staged_gaussians_per_step"])
PY
done
timeout 600 python tools/bench_configs.py --steps 20 --cams5 8 --only 3 > gpurun_out/configs.jsonl 2> gpurun_out/configs.err
python - gpurun_out/configs.jsonl <<'PY'
import json,sys
for l in open(sys.argv[1]):
    d=json.loads(l); print("  ", d["config"][:40], d["ms_per_step"], d["stages_ms"])
PY
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"bg_reach|merge_kernel" -c 40 --csv --log-file gpurun_out/launches_reach.csv $CMD > gpurun_out/ncu_launches.log 2>&1
python tools/launch_summary.py gpurun_out/launches_reach.csv
