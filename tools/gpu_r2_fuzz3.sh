#!/bin/bash
# the ablation knobs of the compacted lists combined with the view-bucket records
mkdir -p gpurun_out
SAS_COMPACT_RATIO=1 SAS_COMPACT_MAX=20 timeout 900 python tools/fuzz_parity.py 150 32000 > gpurun_out/fuzz_compact_gate.log 2>&1; echo "RATIO=1 MAX=20: exit $? $(tail -1 gpurun_out/fuzz_compact_gate.log | cut -c1-200)"
SAS_COMPACT_RATIO=1 SAS_COMPACT_MAX=0 timeout 900 python tools/fuzz_parity.py 60 37000 > gpurun_out/fuzz_compact_gate0.log 2>&1; echo "RATIO=1 MAX=0: exit $? $(tail -1 gpurun_out/fuzz_compact_gate0.log | cut -c1-200)"
SAS_COMPACT_RATIO=1 SAS_VARIANT=0x12cf timeout 900 python tools/fuzz_parity.py 100 38000 > gpurun_out/fuzz_bit12.log 2>&1; echo "RATIO=1 0x12cf: exit $? $(tail -1 gpurun_out/fuzz_bit12.log | cut -c1-200)"
SAS_COMPACT_RATIO=1 SAS_VARIANT=0x36cf timeout 900 python tools/fuzz_parity.py 40 39000 > gpurun_out/fuzz_split13.log 2>&1; echo "RATIO=1 0x36cf (split + bit 13 requested): exit $? $(tail -1 gpurun_out/fuzz_split13.log | cut -c1-200)"
python - <<'PY'
import ast, glob
for f in sorted(glob.glob("gpurun_out/fuzz_*.log")):
    rows = [ast.literal_eval(l) for l in open(f) if l.startswith("{")]
    print(f, len(rows), "cases", "img_bitdiff total", sum(r["img_bitdiff"] for r in rows), "max M", max(r["M"] for r in rows), "max list", max(r["max_list"] for r in rows))
PY
