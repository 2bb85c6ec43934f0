mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sparse.py -m gpu -q -x -k "narrow or single_conv or resnet3d_small" 2>&1 | tail -15 > gpurun_out/r2_t20.log
for d in 5 3 9; do
P3D_WM_D=$d timeout 300 python bench.py --no-cpu-baseline --no-second-geometry --steps 100 --warmup 10 > gpurun_out/r2_b20_d$d.json 2> gpurun_out/r2_b20_d$d.err
done
timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv_wm --launch-skip 3 --launch-count 3 -o gpurun_out/r02_wm4 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-second-geometry > /dev/null 2> gpurun_out/r2_ncu20.err
echo done
