mkdir -p gpurun_out
timeout 300 python bench.py --no-cpu-baseline --no-second-geometry --steps 300 --warmup 20 --in-flight 6 > gpurun_out/r2_b39_L6.json 2> gpurun_out/r2_b39_L6.err
timeout 300 python bench.py --no-cpu-baseline --no-second-geometry --steps 300 --warmup 20 --in-flight 4 > gpurun_out/r2_b39_L4.json 2> gpurun_out/r2_b39_L4.err
echo done
