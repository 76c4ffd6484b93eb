#!/bin/bash
# last sweep on the final tree: default (bits 12 + 13), the path forced on everywhere, and the per-part-list ablation (bit 14)
mkdir -p gpurun_out
timeout 900 python tools/fuzz_parity.py 250 40000 > gpurun_out/fuzz_default.log 2>&1; echo "default: exit $? $(tail -1 gpurun_out/fuzz_default.log | cut -c1-200)"
SAS_COMPACT_RATIO=1 timeout 900 python tools/fuzz_parity.py 150 41000 > gpurun_out/fuzz_ratio1.log 2>&1; echo "RATIO=1: exit $? $(tail -1 gpurun_out/fuzz_ratio1.log | cut -c1-200)"
SAS_COMPACT_RATIO=1 SAS_VARIANT=0x72cf timeout 900 python tools/fuzz_parity.py 150 42000 > gpurun_out/fuzz_parts.log 2>&1; echo "RATIO=1 0x72cf: exit $? $(tail -1 gpurun_out/fuzz_parts.log | cut -c1-200)"
python - <<'PY'
import ast, glob
for f in sorted(glob.glob("gpurun_out/fuzz_*.log")):
    rows = [ast.literal_eval(l) for l in open(f) if l.startswith("{")]
    print(f, len(rows), "cases", "img_bitdiff total", sum(r["img_bitdiff"] for r in rows), "max M", max(r["M"] for r in rows), "max list", max(r["max_list"] for r in rows))
PY
