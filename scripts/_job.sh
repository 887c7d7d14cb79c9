mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/gpu_tests.log
timeout 300 python scripts/ncu_target.py 16 64 1 148 > gpurun_out/ncu_plain.log 2>&1; echo "plain rc=$?"; tail -1 gpurun_out/ncu_plain.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_right_compress -s 2 -c 1 -f -o gpurun_out/r1_fused_L16_chi64_B148 python scripts/ncu_target.py 16 64 1 148 > gpurun_out/ncu_full.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out/*.ncu-rep
