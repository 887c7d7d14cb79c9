mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; echo "bench rc=$?"
python -c "
import json; j=json.load(open('gpurun_out/bench_r1.json')); r=j['roofline']; print(j['value'], j['e2e']['value'], r['peak'], r.get('fp64_tensor_peak_dmma_probe'), r['frac'], j['cpu_baseline']['value'])"
