#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?: $(tail -1 gpurun_out/pytest_gpu.log)"
timeout 600 python tools/fuzz_parity.py ${FUZZ:-80} 13000 > gpurun_out/fuzz.log 2>&1
echo "fuzz exit $?: $(tail -1 gpurun_out/fuzz.log)"
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], round(d["value"]), "fps", round(d["ms_per_step"],3), "ms/step; serial", round(d["serial"]["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), {k:v["ms"] for k,v in d["stages"].items()}, d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
    for o in d.get("other_configs",[]): print("   ", o["config"][:30], o["ms_per_step"], o.get("stages_ms"))
except Exception as e:
    print(sys.argv[2], "no bench line", e)
PY
}
timeout 600 python bench.py --steps 100 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; show gpurun_out/bench.log default
for v in ${PVARS:-1 3}; do
  SAS_PROJECT=$v timeout 300 python bench.py --steps 100 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/bench_p$v.log 2> gpurun_out/bench_p$v.err; show gpurun_out/bench_p$v.log SAS_PROJECT=$v
done
