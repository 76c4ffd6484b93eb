#!/bin/bash
# first GPU call: parity tests, smoke, bench, then (only if the plain run exited 0) ncu launch list + full captures
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/bench.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:composite_kernel -s 3 -c 2 -o gpurun_out/prof_composite $CMD > gpurun_out/ncu_composite.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:project_kernel -s 3 -c 1 -o gpurun_out/prof_project $CMD > gpurun_out/ncu_project.log 2>&1
echo "ncu chain exit $?" > gpurun_out/ncu_exit.txt
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -3; cat gpurun_out/bench.log | cut -c1-600; tail -3 gpurun_out/bench.err
