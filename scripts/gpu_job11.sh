timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; tail -n 2 gpurun_out/bench_r1.err; cat gpurun_out/bench_r1.json
for args in "--workload heisenberg --chi 32 --layers 6 --batch 512" "--L 50 --chi 16 --layers 4 --batch 1024" "--L 100 --chi 32 --layers 6 --batch 512"; do
n=$(echo $args | tr -d ' -' | cut -c1-24)
timeout 600 python bench.py $args > gpurun_out/bench_$n.json 2>/dev/null; python -c "import json,sys; d=json.load(open('gpurun_out/bench_$n.json')); print(d['config']['workload'], round(d['value'],1), 'MPS/s e2e', round(d['e2e']['value'],1), 'cpu', round(d['cpu_baseline']['value'],2))"
done
