#!/bin/bash
# bin sort: parity, A/B against the radix sort, other configs; then compact ncu captures (64 MiB limit on gpurun_out)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?: $(tail -1 gpurun_out/pytest_gpu.log)"
[ "$(tail -1 gpurun_out/pytest_gpu.log | grep -c failed)" != "0" ] && tail -60 gpurun_out/pytest_gpu.log
timeout 900 python tools/fuzz_parity.py 100 16000 > gpurun_out/fuzz.log 2>&1
echo "fuzz exit $?: $(tail -1 gpurun_out/fuzz.log)"
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], round(d["value"]), "fps", round(d["ms_per_step"],3), "ms/step; serial", round(d["serial"]["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), {k:v["ms"] for k,v in d["stages"].items()}, d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print(sys.argv[2], "no bench line", e)
PY
}
for v in bin radix; do
  SAS_SORT=$v timeout 300 python bench.py --steps 100 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/bench_s$v.log 2> gpurun_out/bench_s$v.err; show gpurun_out/bench_s$v.log SAS_SORT=$v
  SAS_SORT=$v timeout 600 python tools/bench_configs.py --steps 10 --cams5 8 > gpurun_out/configs_s$v.jsonl 2> gpurun_out/configs_s$v.err
  python - gpurun_out/configs_s$v.jsonl <<'PY'
import json,sys
for l in open(sys.argv[1]):
    d=json.loads(l); print("   ", d["config"][:40], d["ms_per_step"], d["stages_ms"])
PY
done
if [ -z "$NO_NCU" ]; then
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none -s 60 -c 26 -f -o gpurun_out/prof_step $CMD > gpurun_out/ncu_step.log 2>&1
echo "ncu step exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:composite_kernel -s 6 -c 1 -f -o gpurun_out/prof_composite $CMD > gpurun_out/ncu_comp.log 2>&1
echo "ncu composite exit $?"
SAS_VARIANT=0xacf timeout 900 ncu --set full --clock-control none --import-source on -k regex:composite_kernel -s 6 -c 1 -f -o gpurun_out/prof_bulk_composite $CMD > gpurun_out/ncu_bulk.log 2>&1
echo "ncu bulk exit $?"
ls -la gpurun_out/*.ncu-rep
fi
