mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sparse.py -m gpu -q -x -k "narrow or single_conv" 2>&1 | tail -15 > gpurun_out/r2_t17.log
timeout 300 python bench.py --no-cpu-baseline --no-second-geometry --steps 100 --warmup 10 > gpurun_out/r2_b17_wm.json 2> gpurun_out/r2_b17_wm.err
P3D_SPARSE_WM=0 timeout 300 python bench.py --no-cpu-baseline --no-second-geometry --steps 100 --warmup 10 > gpurun_out/r2_b17_tc.json 2> gpurun_out/r2_b17_tc.err
echo done
