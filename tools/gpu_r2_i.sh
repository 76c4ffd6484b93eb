#!/bin/bash
# bulk-async staging variant (bit 11): parity + A/B; then the all-kernel ncu capture of one bench step
mkdir -p gpurun_out
SAS_VARIANT=0xacf timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_fuzz_gpu.py tests/test_compat_gpu.py -m gpu -q --timeout=600 -x > gpurun_out/pytest_bulk.log 2>&1
echo "pytest(bulk) exit $?: $(tail -1 gpurun_out/pytest_bulk.log)"
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], round(d["value"]), "fps", round(d["ms_per_step"],3), "ms/step; serial", round(d["serial"]["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), {k:v["ms"] for k,v in d["stages"].items()}, d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print(sys.argv[2], "no bench line", e)
PY
}
for v in 0x2cf 0xacf 0x2cf 0xacf; do
  SAS_VARIANT=$v timeout 300 python bench.py --steps 100 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/bench_v$v.log 2> gpurun_out/bench_v$v.err; show gpurun_out/bench_v$v.log SAS_VARIANT=$v
done
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -s 60 -c 60 -f -o gpurun_out/prof_step $CMD > gpurun_out/ncu_step.log 2>&1
echo "ncu step exit $?"
SAS_VARIANT=0xacf timeout 300 $CMD > gpurun_out/plain_bulk.log 2>&1 && \
SAS_VARIANT=0xacf timeout 900 ncu --set full --clock-control none --import-source on -k regex:composite_kernel -s 6 -c 1 -f -o gpurun_out/prof_bulk_composite $CMD > gpurun_out/ncu_bulk.log 2>&1
echo "ncu bulk exit $?"
