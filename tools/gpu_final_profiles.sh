#!/bin/bash
# gpurun payload: bench line, ncu launch list of the same command, ncu --set full captures of the top kernels
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"
timeout 300 python bench.py --impl reference --steps 10 --warmup 1 > gpurun_out/bench_ref.log 2>/dev/null; echo "ref exit $?"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launch list exit $?"
for spec in ${PROF_SPECS:-composite_kernel:4:1 project_kernel:8:2 tile_sort_kernel:20:5 merge_kernel:4:1 emit_kernel:8:2}; do
  name=${spec%%:*}; rest=${spec#*:}; skip=${rest%%:*}; cnt=${rest#*:}
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$name -s $skip -c $cnt -f -o gpurun_out/prof_$name $CMD > gpurun_out/ncu_$name.log 2>&1
  echo "ncu $name exit $?"
done
