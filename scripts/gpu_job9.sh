timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -2
for args in "--workload heisenberg --chi 32 --layers 6 --batch 512" "--L 50 --chi 16 --layers 4 --batch 1024" "--L 100 --chi 32 --layers 6 --batch 512" ""; do
timeout 600 python bench.py $args --steps 2 --warmup 1 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['config']['workload'], round(d['value'],1), 'MPS/s', round(d['ms_per_step'],1),'ms/step', d['roofline']['layer_kernel_ms'])"
done
