mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_voxelize.py tests/test_gpu_pipeline.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r2_t22.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:vox_ -c 24 --csv --log-file gpurun_out/r2_vox_launches.csv python tools/op_bench.py --only voxelize --c3-only --iters 2 > /dev/null 2>&1
timeout 300 python bench.py --no-cpu-baseline --no-second-geometry --steps 100 --warmup 10 > gpurun_out/r2_b22.json 2> gpurun_out/r2_b22.err
echo done
