mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2_b40.json 2> gpurun_out/r2_b40.err
echo done
