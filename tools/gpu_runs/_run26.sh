mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_voxelize.py tests/test_gpu_pipeline.py tests/test_gpu_sparse.py -m gpu -q 2>&1 | tail -8 > gpurun_out/r2_t26.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:vox_ -c 28 --csv --log-file gpurun_out/r2_vox_launches.csv python tools/op_bench.py --only voxelize --iters 2 > /dev/null 2>&1
timeout 300 python tools/op_bench.py --only voxelize > gpurun_out/r2_op26.jsonl 2>&1
timeout 300 python bench.py --no-cpu-baseline --no-second-geometry --steps 100 --warmup 10 > gpurun_out/r2_b26.json 2> gpurun_out/r2_b26.err
P3D_SIDE_STREAM=1 timeout 300 python bench.py --no-cpu-baseline --no-second-geometry --steps 100 --warmup 10 > gpurun_out/r2_b26_side.json 2> gpurun_out/r2_b26_side.err
echo done
