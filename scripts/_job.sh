mkdir -p gpurun_out
PROBE_ENC=1 timeout 600 python scripts/probe.py 100 64 4 1184 2 2>&1 | tail -2
PROBE_ENC=1 timeout 600 python scripts/probe.py 100 64 4 740 2 2>&1 | tail -2
timeout 600 python bench.py --L 50 --chi 16 --layers 4 --batch 1024 > gpurun_out/cfg2.json 2>gpurun_out/cfg2.err; tail -c 400 gpurun_out/cfg2.json | head -c 100; echo
timeout 600 python bench.py --workload heisenberg --chi 32 --layers 6 --batch 512 > gpurun_out/cfg3a.json 2>gpurun_out/cfg3a.err
timeout 600 python bench.py --workload aklt --L 100 --chi 32 --layers 6 --batch 512 > gpurun_out/cfg3b.json 2>gpurun_out/cfg3b.err
timeout 600 python bench.py --L 100 --chi 32 --layers 6 --batch 512 > gpurun_out/cfg3c.json 2>gpurun_out/cfg3c.err
for f in cfg2 cfg3a cfg3b cfg3c; do python -c "
import json; j=json.load(open('gpurun_out/$f.json')); print('$f', j['value'], j['e2e']['value'], j['cpu_baseline']['value'], j['roofline']['frac'])"; done
