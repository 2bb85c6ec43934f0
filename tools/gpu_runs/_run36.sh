mkdir -p gpurun_out
timeout 300 python bench.py --no-cpu-baseline --no-second-geometry --steps 200 --warmup 10 --in-flight 1 > gpurun_out/r2_b36_h1000_L1.json 2> gpurun_out/r2_b36_h1000_L1.err
timeout 300 python bench.py --no-cpu-baseline --no-second-geometry --steps 200 --warmup 10 > gpurun_out/r2_b36_h1000_L3.json 2> gpurun_out/r2_b36_h1000_L3.err
echo done
