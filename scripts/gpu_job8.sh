set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; tail -n 3 gpurun_out/bench_r1.err; cat gpurun_out/bench_r1.json
CMD="python bench.py --batch 148 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 600 $CMD > gpurun_out/plain_launches.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file gpurun_out/launches_r1.csv $CMD > gpurun_out/ncu_launches.log 2>&1
tail -n 1 gpurun_out/ncu_launches.log | cut -c1-150
