CMD="python scripts/ncu_target.py 16 64 1 148"
$CMD > gpurun_out/plain10.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_right_compress -s 2 -c 1 -o gpurun_out/prof_rcompress_r1c $CMD > gpurun_out/ncu10.log 2>&1
tail -n 2 gpurun_out/ncu10.log
