mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_raw.py tests/test_gpu_sparse.py tests/test_gpu_pipeline.py -m gpu -q 2>&1 | tail -12 > gpurun_out/r2_t34.log
timeout 300 python bench.py --no-cpu-baseline --no-second-geometry --steps 200 --warmup 10 > gpurun_out/r2_b34.json 2> gpurun_out/r2_b34.err
echo done
