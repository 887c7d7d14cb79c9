timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "facade" 2>&1 | tail -2
timeout 300 python scripts/readme_latency.py 2>&1 | tail -2
