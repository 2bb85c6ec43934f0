timeout 600 python -m pytest tests/test_gpu_pipeline.py -m gpu -q -k "fused or dense_head" 2>&1 | tail -12 > gpurun_out/r2_t13.log
echo done
