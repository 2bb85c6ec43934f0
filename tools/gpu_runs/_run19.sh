mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sparse.py -m gpu -q -x -k "narrow or single_conv or resnet3d_small or sparsenet3d" 2>&1 | tail -15 > gpurun_out/r2_t19.log
timeout 300 python bench.py --no-cpu-baseline --no-second-geometry --steps 100 --warmup 10 > gpurun_out/r2_b19_wm.json 2> gpurun_out/r2_b19_wm.err
P3D_WM_D=4 timeout 300 python bench.py --no-cpu-baseline --no-second-geometry --steps 100 --warmup 10 > gpurun_out/r2_b19_wm_d4.json 2> gpurun_out/r2_b19_wm_d4.err
timeout 300 python tools/op_bench.py --only voxelize > gpurun_out/r2_op19.jsonl 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv_wm --launch-skip 3 --launch-count 3 -o gpurun_out/r02_wm3 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-second-geometry > /dev/null 2> gpurun_out/r2_ncu19.err
echo done
