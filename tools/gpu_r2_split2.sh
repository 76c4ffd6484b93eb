#!/bin/bash
mkdir -p gpurun_out
for v in 0x2cf 0x6cf; do
  SAS_VARIANT=$v timeout 600 python tools/bench_configs.py --steps 10 --cams5 4 > gpurun_out/configs_v$v.jsonl 2> gpurun_out/configs_v$v.err
  echo "configs $v exit $?"; python - gpurun_out/configs_v$v.jsonl <<'PY'
import json,sys
for l in open(sys.argv[1]):
    d=json.loads(l); print(d["config"][:40], d["ms_per_step"], d["stages_ms"])
PY
done
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
export SAS_VARIANT=0x6cf
timeout 300 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:composite_kernel -s 6 -c 1 -f -o gpurun_out/prof_split_composite $CMD > gpurun_out/ncu_split.log 2>&1
echo "ncu exit $?"
