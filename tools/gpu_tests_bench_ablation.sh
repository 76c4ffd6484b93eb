#!/bin/bash
# parity tests, smoke, bench (default + SAS_VARIANT ablations), ncu launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?: $(tail -1 gpurun_out/pytest_gpu.log)"
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?: $(tail -1 gpurun_out/smoke.log)"
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], round(d["value"]), "fps", round(d["ms_per_step"],3), "ms/step; e2e", round(d["e2e"]["value"]), {k:v["ms"] for k,v in d["stages"].items()}, d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print(sys.argv[2], "no bench line", e)
PY
}
timeout 600 python bench.py --steps 50 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"; show gpurun_out/bench.log default
for v in ${VARIANTS:-0xfe 0xfd 0xfb 0xf7 0x00}; do
  SAS_VARIANT=$v timeout 300 python bench.py --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v$v.log 2> gpurun_out/bench_v$v.err; show gpurun_out/bench_v$v.log SAS_VARIANT=$v
done
if [ -z "$NO_NCU" ]; then
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu chain exit $?"
fi
