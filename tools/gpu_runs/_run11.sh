timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-second-geometry > gpurun_out/r2_b11_underncu.json 2> gpurun_out/r2_b11.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'dense_conv_f16|conv_f16_kernel' --launch-skip 130 --launch-count 40 -o gpurun_out/r02_tc_kernels python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-second-geometry > /dev/null 2> gpurun_out/r2_b11b.err
echo done
