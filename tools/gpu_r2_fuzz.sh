#!/bin/bash
# parity evidence for the round-2 build: randomised sweep against the oracle, default build and every ablation variant
mkdir -p gpurun_out
timeout 1500 python tools/fuzz_parity.py 500 20000 > gpurun_out/fuzz_default.log 2>&1; echo "default: exit $? $(tail -1 gpurun_out/fuzz_default.log)"
SAS_VARIANT=0x6cf timeout 600 python tools/fuzz_parity.py 100 21000 > gpurun_out/fuzz_split.log 2>&1; echo "split 0x6cf: exit $? $(tail -1 gpurun_out/fuzz_split.log)"
SAS_VARIANT=0xacf timeout 600 python tools/fuzz_parity.py 100 22000 > gpurun_out/fuzz_bulk.log 2>&1; echo "bulk 0xacf: exit $? $(tail -1 gpurun_out/fuzz_bulk.log)"
SAS_SORT=bin timeout 600 python tools/fuzz_parity.py 100 23000 > gpurun_out/fuzz_binsort.log 2>&1; echo "SAS_SORT=bin: exit $? $(tail -1 gpurun_out/fuzz_binsort.log)"
SAS_MORTON=1 timeout 600 python tools/fuzz_parity.py 100 24000 > gpurun_out/fuzz_morton.log 2>&1; echo "SAS_MORTON=1: exit $? $(tail -1 gpurun_out/fuzz_morton.log)"
SAS_SORT_BIG=0 timeout 600 python tools/fuzz_parity.py 100 25000 > gpurun_out/fuzz_nobig.log 2>&1; echo "SAS_SORT_BIG=0: exit $? $(tail -1 gpurun_out/fuzz_nobig.log)"
python - <<'PY'
import ast, glob
for f in sorted(glob.glob("gpurun_out/fuzz_*.log")):
    rows = [ast.literal_eval(l) for l in open(f) if l.startswith("{")]
    print(f, len(rows), "cases", "img_bitdiff total", sum(r["img_bitdiff"] for r in rows), "max M", max(r["M"] for r in rows), "max list", max(r["max_list"] for r in rows))
PY
