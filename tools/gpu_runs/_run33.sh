mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 10 --no-cpu-baseline --no-second-geometry > gpurun_out/r2_b33_n2.json 2> gpurun_out/r2_b33_n2.err
echo done
