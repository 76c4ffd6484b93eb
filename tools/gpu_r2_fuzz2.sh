#!/bin/bash
# parity evidence for the last round-2 build (compacted compositor lists, streaming merge): randomised sweep against the oracle
mkdir -p gpurun_out
timeout 1500 python tools/fuzz_parity.py 300 30000 > gpurun_out/fuzz_default.log 2>&1; echo "default: exit $? $(tail -1 gpurun_out/fuzz_default.log)"
SAS_COMPACT_RATIO=1 timeout 1500 python tools/fuzz_parity.py 300 31000 > gpurun_out/fuzz_compact_all.log 2>&1; echo "SAS_COMPACT_RATIO=1: exit $? $(tail -1 gpurun_out/fuzz_compact_all.log)"
SAS_COMPACT_RATIO=1 SAS_COMPACT_MAX=20 timeout 900 python tools/fuzz_parity.py 100 32000 > gpurun_out/fuzz_compact_gate.log 2>&1; echo "SAS_COMPACT_RATIO=1 SAS_COMPACT_MAX=20: exit $? $(tail -1 gpurun_out/fuzz_compact_gate.log)"
SAS_VARIANT=0x2cf timeout 900 python tools/fuzz_parity.py 100 33000 > gpurun_out/fuzz_nocompact.log 2>&1; echo "SAS_VARIANT=0x2cf: exit $? $(tail -1 gpurun_out/fuzz_nocompact.log)"
python - <<'PY'
import ast, glob
for f in sorted(glob.glob("gpurun_out/fuzz_*.log")):
    rows = [ast.literal_eval(l) for l in open(f) if l.startswith("{")]
    print(f, len(rows), "cases", "img_bitdiff total", sum(r["img_bitdiff"] for r in rows), "max M", max(r["M"] for r in rows), "max list", max(r["max_list"] for r in rows))
PY
