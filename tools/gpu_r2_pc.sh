#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs"
for v in 0x72cf 0x32cf; do
SAS_VARIANT=$v timeout 600 ncu --metrics gpu__time_duration.sum,sm__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,lts__t_bytes.sum --clock-control none -k regex:"merge_kernel|composite_kernel" -s 8 -c 4 --csv --log-file gpurun_out/m_$v.csv $CMD > gpurun_out/ncu_$v.log 2>&1
python - gpurun_out/m_$v.csv $v <<'PY'
import csv,sys
rows=[r for r in csv.DictReader(l for l in open(sys.argv[1]) if not l.startswith("=="))]
agg={}
for r in rows:
    k=(r["Kernel Name"][:28], r["Metric Name"])
    agg.setdefault(k,[]).append(float(r["Metric Value"].replace(",","")))
for k,v in sorted(agg.items()): print(sys.argv[2], k, sum(v)/len(v), r["Metric Unit"] if False else "")
PY
done
