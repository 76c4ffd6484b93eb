#!/bin/bash
# A/B the compositor staging pipelines (SAS_COMPOSITE=0..3): parity subset + bench for each
mkdir -p gpurun_out
for v in 0 1 2 3; do
  SAS_COMPOSITE=$v timeout 600 python -m pytest tests -m gpu -q --timeout=300 -k "small_scene or config2 or bucket_sort or views_shared or full_size" > gpurun_out/ab_pytest_$v.log 2>&1
  echo "variant $v pytest exit $?: $(tail -1 gpurun_out/ab_pytest_$v.log)"
  SAS_COMPOSITE=$v timeout 300 python bench.py --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/ab_bench_$v.log 2> gpurun_out/ab_bench_$v.err
  echo "variant $v bench exit $?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/ab_bench_$v.log').read().strip().splitlines()[-1])
    print("variant $v", round(d["value"]), "fps", round(d["ms_per_step"],3), "ms/step; e2e", round(d["e2e"]["value"]), {k:v["ms"] for k,v in d["stages"].items()})
except Exception as e:
    print("variant $v: no bench line", e)
PY
done
