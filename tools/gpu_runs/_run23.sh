mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pipeline.py -m gpu -q -x -k fused_pixel 2>&1 | tail -60 > gpurun_out/r2_t23.log
P3D_SPARSE_WM=0 timeout 600 python -m pytest tests/test_gpu_pipeline.py -m gpu -q -x -k fused_pixel 2>&1 | tail -8 > gpurun_out/r2_t23_tc.log
P3D_PDL=0 timeout 600 python -m pytest tests/test_gpu_pipeline.py -m gpu -q -x -k fused_pixel 2>&1 | tail -8 > gpurun_out/r2_t23_nopdl.log
echo done
