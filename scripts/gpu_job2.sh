set -x
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -5
timeout 300 python scripts/probe.py 100 64 4 296 1 2>&1 | tail -2
timeout 300 python scripts/probe.py 50 16 4 1024 2 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -12 gpurun_out/bench_default.err; cat gpurun_out/bench_default.json
