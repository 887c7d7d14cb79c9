mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/gpu_tests.log
timeout 900 python bench.py > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; echo "bench rc=$?"
timeout 600 python bench.py --L 50 --chi 16 --layers 4 --batch 1024 > gpurun_out/cfg2.json 2>gpurun_out/cfg2.err
timeout 600 python bench.py --workload heisenberg --chi 32 --layers 6 --batch 512 > gpurun_out/cfg3a.json 2>gpurun_out/cfg3a.err
timeout 600 python bench.py --workload aklt --L 100 --chi 32 --layers 6 --batch 512 > gpurun_out/cfg3b.json 2>gpurun_out/cfg3b.err
timeout 600 python bench.py --L 100 --chi 32 --layers 6 --batch 512 > gpurun_out/cfg3c.json 2>gpurun_out/cfg3c.err
for f in bench_r1 cfg2 cfg3a cfg3b cfg3c; do python -c "
import json; j=json.load(open('gpurun_out/$f.json')); print('$f', round(j['value'],2), round(j['e2e']['value'],2), round(j['cpu_baseline']['value'],2), round(j['roofline']['frac'],3), j['gpu_launches'])"; done
timeout 300 python scripts/readme_latency.py 2>&1 | tail -2
