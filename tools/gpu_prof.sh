#!/bin/bash
# ncu --set full captures of selected kernels (one plain run first; all ncu runs count as one use)
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
for spec in $PROF_SPECS; do
  name=${spec%%:*}; rest=${spec#*:}; skip=${rest%%:*}; cnt=${rest#*:}
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$name -s $skip -c $cnt -f -o gpurun_out/prof_$name $CMD > gpurun_out/ncu_$name.log 2>&1
  echo "ncu $name exit $?"
done
