set -x
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -3 gpurun_out/bench_default.err; cat gpurun_out/bench_default.json
python -m pytest tests -m gpu -q -k "p4 or heisenberg" 2>&1 | tail -3
CMD="python bench.py --batch 148 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
tail -2 gpurun_out/ncu1.log
$CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_right_compress -s 1 -c 1 -o gpurun_out/prof_rcompress $CMD > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log
ls -la gpurun_out
