set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/gpu_tests.log 2>&1; echo "tests rc=$?"
tail -15 gpurun_out/gpu_tests.log
timeout 300 python scripts/probe.py 24 128 1 148 1 > gpurun_out/probe128.log 2>&1; tail -3 gpurun_out/probe128.log
timeout 300 python scripts/probe.py 20 256 1 32 1 > gpurun_out/probe256.log 2>&1; tail -3 gpurun_out/probe256.log
