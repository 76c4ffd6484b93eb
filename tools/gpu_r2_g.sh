#!/bin/bash
# launch list of configs[4] (2 cameras) and configs[3] shard: where does sort / emit time go
mkdir -p gpurun_out
for c in 4 3; do
CMD="python tools/bench_configs.py --steps 4 --cams5 2 --only $c"
SAS_GRAPH=0 timeout 600 $CMD > gpurun_out/plain_c$c.log 2>&1 && \
SAS_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_c$c.csv $CMD > gpurun_out/ncu_c$c.log 2>&1
echo "ncu c$c exit $?"
python tools/launch_summary.py gpurun_out/launches_c$c.csv | head -30
done
