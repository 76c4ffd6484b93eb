#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_api_r2_gpu.py -m gpu -q --timeout=600 -x --durations=10 > gpurun_out/pytest_api.log 2>&1
echo "pytest(api) exit $?: $(tail -1 gpurun_out/pytest_api.log)"
timeout 1500 python -m pytest tests/test_configs_gpu.py -m gpu -q --timeout=1200 -x --durations=10 -s > gpurun_out/pytest_configs.log 2>&1
echo "pytest(configs) exit $?: $(tail -1 gpurun_out/pytest_configs.log)"
timeout 900 python -m pytest tests -m gpu -q --timeout=600 -x --deselect tests/test_configs_gpu.py > gpurun_out/pytest_rest.log 2>&1
echo "pytest(rest) exit $?: $(tail -1 gpurun_out/pytest_rest.log)"
bash tools/gpu_r2_split2.sh
