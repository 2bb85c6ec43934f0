mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sparse.py tests/test_gpu_voxelize.py tests/test_gpu_dense.py -m gpu -q 2>&1 | tail -5 > gpurun_out/r2_t35.log
timeout 300 python tools/op_bench.py --only scatter,c2 > gpurun_out/r2_op35.jsonl 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:scat_ -c 12 --csv --log-file gpurun_out/r2_scat_launches.csv python tools/op_bench.py --only scatter --iters 2 > /dev/null 2>&1
echo done
