#!/bin/bash
# round 2: split compositor (SAS_VARIANT bit 10) -- parity under the variant, then bench A/B
mkdir -p gpurun_out
V=${V:-0x6cf}
SAS_VARIANT=$V timeout 900 python -m pytest tests -m gpu -q --timeout=600 -x > gpurun_out/pytest_split.log 2>&1
echo "pytest(split) exit $?: $(tail -1 gpurun_out/pytest_split.log)"
SAS_VARIANT=$V timeout 600 python tools/fuzz_parity.py ${FUZZ:-60} 12000 > gpurun_out/fuzz_split.log 2>&1
echo "fuzz(split) exit $?: $(tail -1 gpurun_out/fuzz_split.log)"
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], round(d["value"]), "fps", round(d["ms_per_step"],3), "ms/step; serial", round(d["serial"]["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), {k:v["ms"] for k,v in d["stages"].items()}, d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print(sys.argv[2], "no bench line", e)
PY
}
for v in ${VARIANTS:-0x2cf 0x6cf}; do
  SAS_VARIANT=$v timeout 300 python bench.py --steps 100 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v$v.log 2> gpurun_out/bench_v$v.err; show gpurun_out/bench_v$v.log SAS_VARIANT=$v
done
