mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none -k regex:'dense_conv_f16_kernel|head_out9_kernel' --launch-skip 34 --launch-count 17 -o gpurun_out/r02_dense python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-second-geometry --in-flight 1 > /dev/null 2> gpurun_out/r2_ncu38.err
timeout 300 python bench.py --no-cpu-baseline --no-second-geometry --steps 100 --warmup 10 > gpurun_out/r2_b38.json 2> gpurun_out/r2_b38.err
echo done
