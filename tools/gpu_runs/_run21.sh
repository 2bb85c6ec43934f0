mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sparse.py -m gpu -q -x -k "narrow or single_conv or resnet3d_small" 2>&1 | tail -5 > gpurun_out/r2_t21.log
for d in 0 3; do
P3D_WM_D=$d timeout 300 python bench.py --no-cpu-baseline --no-second-geometry --steps 100 --warmup 10 > gpurun_out/r2_b21_d$d.json 2> gpurun_out/r2_b21_d$d.err
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:vox_ -c 40 --csv --log-file gpurun_out/r2_vox_launches.csv python tools/op_bench.py --only voxelize --c3-only --iters 2 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches21.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-second-geometry > gpurun_out/r2_b21_ncu.log 2>&1
echo done
