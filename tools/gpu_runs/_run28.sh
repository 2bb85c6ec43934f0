mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_voxelize.py -m gpu -q 2>&1 | tail -4 > gpurun_out/r2_t28.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:vox_ -c 40 --csv --log-file gpurun_out/r2_vox_launches.csv python tools/op_bench.py --only voxelize --c3-only --iters 2 > /dev/null 2>&1
timeout 300 python tools/op_bench.py --only voxelize > gpurun_out/r2_op28.jsonl 2>&1
echo done
