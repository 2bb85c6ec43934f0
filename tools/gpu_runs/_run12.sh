timeout 600 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_sparse.py tests/test_gpu_boxes.py -m gpu -q 2>&1 | tail -12 > gpurun_out/r2_t12.log
timeout 600 python bench.py --steps 100 --no-second-geometry > gpurun_out/r2_b12.json 2> gpurun_out/r2_b12.err
echo done
