mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r2_t37_all.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke37.log 2>&1
timeout 900 python bench.py > gpurun_out/r2_b37.json 2> gpurun_out/r2_b37.err
timeout 300 python bench.py --no-cpu-baseline --no-second-geometry --steps 200 --warmup 10 --in-flight 4 > gpurun_out/r2_b37_L4.json 2> gpurun_out/r2_b37_L4.err
echo done
