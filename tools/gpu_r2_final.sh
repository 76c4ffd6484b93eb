#!/bin/bash
# final evidence: smoke, tests, bench (+ reference arm), launch list, compact ncu captures
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?: $(tail -1 gpurun_out/smoke.log)"
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?: $(tail -1 gpurun_out/pytest_gpu.log)"
timeout 900 python bench.py --steps 200 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 20 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "bench ref exit $?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
    print("value", round(d["value"]), "serial", round(d["serial"]["value"]), "e2e", round(d["e2e"]["value"]), "e2e_f16", round(d["e2e"]["depth_f16"]["value"]), "blocking", round(d["e2e"]["blocking"]["value"]), "e2e_dev", round(d["e2e_device"]["value"]))
    print({k:v["ms"] for k,v in d["stages"].items()})
    print("strong", d.get("strong_config3"))
    for o in d.get("other_configs",[]): print(o["config"][:40], o["ms_per_step"], o.get("stages_ms"))
    print("cpu", d.get("cpu_baseline"), d.get("cpu_baseline_config0"))
    print("roof", d["roofline"])
    r=json.loads(open('gpurun_out/bench_ref.log').read().strip().splitlines()[-1]); print("ref", r["value"], r["cpu_baseline"]["cores"])
except Exception as e:
    print("no bench line", e); print(open('gpurun_out/bench.err').read()[-3000:])
PY
if [ -z "$NO_NCU" ]; then
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?"
timeout 1500 ncu --set full --clock-control none -s 60 -c 26 -f -o gpurun_out/prof_step $CMD > gpurun_out/ncu_step.log 2>&1
echo "ncu step exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:composite_kernel -s 6 -c 1 -f -o gpurun_out/prof_composite $CMD > gpurun_out/ncu_comp.log 2>&1
echo "ncu composite exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:project_kernel -s 7 -c 1 -f -o gpurun_out/prof_project $CMD > gpurun_out/ncu_proj.log 2>&1
echo "ncu project exit $?"
ls -la gpurun_out/*.ncu-rep
fi
