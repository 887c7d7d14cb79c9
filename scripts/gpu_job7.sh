# dram traffic of the dominant kernel at the bench shape (2 metrics -> few replay passes)
CMD="python bench.py --batch 148 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 600 $CMD > gpurun_out/plain7.log 2>&1 && timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_right_compress -s 1 -c 1 --csv --log-file gpurun_out/traffic_r1.csv $CMD > gpurun_out/ncu7.log 2>&1
grep -E "dram__|duration" gpurun_out/traffic_r1.csv | cut -c1-300
# other BASELINE configs as bench lines (parity cases; not the headline)
timeout 600 python bench.py --L 50 --chi 16 --layers 4 --batch 1024 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; cut -c1-700 gpurun_out/bench_cfg2.json
timeout 600 python bench.py --workload heisenberg --chi 32 --layers 6 --batch 512 > gpurun_out/bench_cfg3a.json 2> gpurun_out/bench_cfg3a.err; tail -n 3 gpurun_out/bench_cfg3a.err; cut -c1-700 gpurun_out/bench_cfg3a.json
timeout 600 python bench.py --L 100 --chi 32 --layers 6 --batch 512 > gpurun_out/bench_cfg3c.json 2> gpurun_out/bench_cfg3c.err; tail -n 3 gpurun_out/bench_cfg3c.err; cut -c1-700 gpurun_out/bench_cfg3c.json
