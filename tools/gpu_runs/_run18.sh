mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_voxelize.py -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r2_t18.log
timeout 300 python tools/op_bench.py --only voxelize > gpurun_out/r2_op18.jsonl 2>&1
P3D_PDL=0 timeout 300 python tools/op_bench.py --only voxelize > gpurun_out/r2_op18_nopdl.jsonl 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv_wm --launch-skip 3 --launch-count 3 -o gpurun_out/r02_wm python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-second-geometry > /dev/null 2> gpurun_out/r2_ncu18.err
echo done
