#!/bin/bash
# bench only, a list of environment settings
mkdir -p gpurun_out
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], round(d["value"]), "fps", round(d["ms_per_step"],3), "ms/step; e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"],3), {k:v["ms"] for k,v in d["stages"].items()}, d["clocks"]["sm_mhz"], d["clocks"]["reasons"], "launches", d["launches_per_step"])
except Exception as e:
    print(sys.argv[2], "no bench line", e)
PY
}
for v in ${ENVS:-X=0}; do
  env $v timeout 300 python bench.py --steps 100 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$v.log 2> gpurun_out/bench_$v.err; show gpurun_out/bench_$v.log $v; tail -2 gpurun_out/bench_$v.err
done
