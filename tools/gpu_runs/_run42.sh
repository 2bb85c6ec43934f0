mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_pipeline.py -m gpu -q -x -k "sweep" 2>&1 | tail -3 > gpurun_out/r2_t42.log
echo done
