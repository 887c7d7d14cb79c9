mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 2 --warmup 3 --no-cpu-baseline 2>gpurun_out/n8.err | grep '^{' > gpurun_out/n8.json; cat gpurun_out/n8.json | head -c 300
