mkdir -p gpurun_out
PROBE_PROF=1 PROBE_ENC=0 timeout 600 python scripts/probe.py 20 256 1 32 1 2>&1 | tail -3
PROBE_PROF=1 PROBE_ENC=0 timeout 600 python scripts/probe.py 24 128 1 148 1 2>&1 | tail -3
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
