#!/bin/bash
# usage: tools/gpu_submit.sh TAG TIMEOUT_S 'command run inside the frozen copy' [gpus]
# Freezes the working tree (with the built .so files) under gpurun_stage/TAG so that editing can go on while the
# call waits for a GPU slot; results land in gpurun_out/TAG/.
TAG=$1; TMO=$2; CMD=$3; GPUS=${4:-1}
ROOT=/root/repo
mkdir -p $ROOT/gpurun_stage
rm -rf $ROOT/gpurun_stage/$TAG; mkdir -p $ROOT/gpurun_stage/$TAG
tar -C $ROOT --exclude=./.git --exclude=./gpurun_out --exclude=./gpurun_stage --exclude='*.ncu-rep' --exclude=.pytest_cache --exclude=__pycache__ -cf - . | tar -C $ROOT/gpurun_stage/$TAG -xf -
GP=""; [ "$GPUS" != "1" ] && GP="--gpus $GPUS"
for try in $(seq 1 40); do
  /usr/local/graft/bin/gpurun $GP --timeout $TMO -- "cd gpurun_stage/$TAG && mkdir -p gpurun_out && ( $CMD ); mkdir -p \$GRAFT_REPO_ROOT/gpurun_out/$TAG && cp -r gpurun_out/. \$GRAFT_REPO_ROOT/gpurun_out/$TAG/" > /tmp/gpu_$TAG.log 2>&1
  rc=$?
  if grep -q "status=transient\|retry in a few minutes\|already running" /tmp/gpu_$TAG.log || [ $rc -eq 3 ]; then sleep 75; continue; fi
  break
done
cat /tmp/gpu_$TAG.log | tail -40
rm -rf $ROOT/gpurun_stage/$TAG
