#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?: $(tail -1 gpurun_out/pytest_gpu.log)"
[ "$(tail -1 gpurun_out/pytest_gpu.log | grep -c failed)" != "0" ] && tail -60 gpurun_out/pytest_gpu.log
timeout 900 python tools/fuzz_parity.py 100 26000 > gpurun_out/fuzz.log 2>&1
echo "fuzz exit $?: $(tail -1 gpurun_out/fuzz.log)"
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], round(d["value"]), "fps", round(d["ms_per_step"],3), "ms/step; serial", round(d["serial"]["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), {k:v["ms"] for k,v in d["stages"].items()}, "staged", d["fp32"]["staged_gaussians_per_step"], "launches", d["launches_per_step"])
except Exception as e:
    print(sys.argv[2], "no bench line", e)
PY
}
for v in ${VARS:-0x12cf 0x2cf 0x12cf 0x2cf}; do
  SAS_VARIANT=$v timeout 300 python bench.py --steps 100 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/bench_v$v.log 2> gpurun_out/bench_v$v.err; show gpurun_out/bench_v$v.log SAS_VARIANT=$v
done
for v in ${VARS2:-0x12cf 0x2cf}; do
  SAS_VARIANT=$v timeout 600 python tools/bench_configs.py --steps 10 --cams5 8 --only 0,2,3 > gpurun_out/configs_v$v.jsonl 2> gpurun_out/configs_v$v.err
  python - gpurun_out/configs_v$v.jsonl $v <<'PY'
import json,sys
for l in open(sys.argv[1]):
    d=json.loads(l); print("   %s"%sys.argv[2], d["config"][:40], d["ms_per_step"], d["stages_ms"])
PY
done
