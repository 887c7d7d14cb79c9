PROBE_ENC=0 PROBE_PROF=1 timeout 300 python scripts/probe.py 100 64 1 148 1 2>&1 | tail -2
PROBE_ENC=0 PROBE_PROF=1 timeout 300 python scripts/probe.py 50 16 1 1024 1 2>&1 | tail -2
CMD="python scripts/ncu_target.py 16 64 1 148"
$CMD > gpurun_out/plain4.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_right_compress -s 2 -c 1 -o gpurun_out/prof_rcompress_r1b $CMD > gpurun_out/ncu4.log 2>&1
tail -n 2 gpurun_out/ncu4.log
