#!/bin/bash
# all GPU tests, config step times, bench with graph on/off
mkdir -p gpurun_out
timeout 2000 python -m pytest tests -m gpu -q --timeout=900 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python tools/bench_configs.py > gpurun_out/bench_configs.log 2> gpurun_out/bench_configs.err; echo "configs exit $?"
python - <<'PY'
import json
for l in open("gpurun_out/bench_configs.log"):
    d=json.loads(l); print(d["config"][:40], d["ms_per_step"], "ms", d["frames_per_s"], "fps", d["stages_ms"])
PY
tail -3 gpurun_out/bench_configs.err
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], round(d["value"]), "fps", round(d["ms_per_step"],3), "ms/step; e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"],3), {k:v["ms"] for k,v in d["stages"].items()}, d["clocks"]["sm_mhz"], d["clocks"]["reasons"], "launches", d["launches_per_step"])
except Exception as e:
    print(sys.argv[2], "no bench line", e)
PY
}
timeout 600 python bench.py --steps 100 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"; show gpurun_out/bench.log default; tail -3 gpurun_out/bench.err
for v in ${ENVS:-SAS_GRAPH=0}; do
  env $v timeout 300 python bench.py --steps 100 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$v.log 2> gpurun_out/bench_$v.err; show gpurun_out/bench_$v.log $v; tail -2 gpurun_out/bench_$v.err
done
