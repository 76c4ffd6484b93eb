#!/bin/bash
# N-GPU bench (weak configs[1] + strong configs[3]); N from $NG
mkdir -p gpurun_out
NG=${NG:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --steps 100 --warmup 3 > gpurun_out/bench_${NG}gpu.log 2> gpurun_out/bench_${NG}gpu.err
echo "bench $NG gpus exit $?"
python - gpurun_out/bench_${NG}gpu.log <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith("{")][-1])
    print("n", d["n_gpus"], "value", round(d["value"]), "serial", round(d["serial"]["value"]), "e2e", round(d["e2e"]["value"]), "e2e_f16", round(d["e2e"]["depth_f16"]["value"]), "blocking", round(d["e2e"]["blocking"]["value"]), "e2e_dev", round(d["e2e_device"]["value"]))
    print("strong", d.get("strong_config3")); print("gather", d.get("obs_gather"))
except Exception as e:
    print("no bench line", e); print(open(sys.argv[1].replace(".log",".err")).read()[-3000:])
PY
