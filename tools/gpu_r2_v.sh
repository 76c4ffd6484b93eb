#!/bin/bash
# view-bucket records with exact reach masks (SAS_VARIANT bit 13): tests, fuzz, A/B against bit 12 alone
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?: $(tail -1 gpurun_out/pytest_gpu.log)"
[ "$(tail -1 gpurun_out/pytest_gpu.log | grep -c failed)" != "0" ] && tail -60 gpurun_out/pytest_gpu.log
timeout 900 python tools/fuzz_parity.py 60 34000 > gpurun_out/fuzz.log 2>&1
echo "fuzz exit $?: $(tail -1 gpurun_out/fuzz.log)"
SAS_COMPACT_RATIO=1 timeout 900 python tools/fuzz_parity.py 100 35000 > gpurun_out/fuzz_r1.log 2>&1
echo "fuzz ratio 1 exit $?: $(tail -1 gpurun_out/fuzz_r1.log)"
for v in ${VARS:-0x72cf 0x32cf 0x72cf 0x32cf}; do
SAS_VARIANT=$v timeout 300 python bench.py --steps 100 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/bench_$v.log 2> gpurun_out/bench_$v.err
python - gpurun_out/bench_$v.log $v <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[2], round(d["value"]), "fps", round(d["ms_per_step"],4), "serial", round(d["serial"]["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), {k:v["ms"] for k,v in d["stages"].items()}, "staged", d["fp32"]["staged_gaussians_per_step"])
PY
done
for v in ${VARS2:-0x72cf 0x32cf}; do
SAS_VARIANT=$v timeout 600 python tools/bench_configs.py --steps 10 --cams5 8 --only 3 > gpurun_out/configs_$v.jsonl 2> gpurun_out/configs_$v.err
python - gpurun_out/configs_$v.jsonl $v <<'PY'
import json,sys
for l in open(sys.argv[1]):
    d=json.loads(l); print("  ", sys.argv[2], d["config"][:40], d["ms_per_step"], d["stages_ms"])
PY
done
