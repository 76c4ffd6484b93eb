#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 100 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
    print("value", round(d["value"]), "serial", round(d["serial"]["value"]), "e2e", round(d["e2e"]["value"]), "e2e_f16", round(d["e2e"]["depth_f16"]["value"]), "blocking", round(d["e2e"]["blocking"]["value"]), "e2e_dev", round(d["e2e_device"]["value"]))
    print({k:v["ms"] for k,v in d["stages"].items()})
    print("strong", d.get("strong_config3"))
    for o in d.get("other_configs",[]): print(o["config"][:30], o["ms_per_step"], o.get("stages_ms"))
    print("cpu", d.get("cpu_baseline"), d.get("cpu_baseline_config0"))
    print("roof", d["roofline"])
except Exception as e:
    print("no bench line", e); print(open('gpurun_out/bench.err').read()[-3000:])
PY
