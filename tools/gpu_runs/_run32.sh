mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_voxelize.py -m gpu -q 2>&1 | tail -3 > gpurun_out/r2_t32.log
timeout 300 python tools/op_bench.py --only voxelize,c2,c4 > gpurun_out/r2_op32.jsonl 2>&1
echo done
