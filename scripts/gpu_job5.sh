set -x
timeout 900 python bench.py > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; tail -n 8 gpurun_out/bench_r1.err; cat gpurun_out/bench_r1.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1_reference.json 2> gpurun_out/bench_r1_reference.err; cat gpurun_out/bench_r1_reference.json; tail -n 3 gpurun_out/bench_r1_reference.err
CMD="python bench.py --batch 148 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 600 $CMD > gpurun_out/plain_launches.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv $CMD > gpurun_out/ncu_launches.log 2>&1
tail -n 2 gpurun_out/ncu_launches.log
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv
nproc
