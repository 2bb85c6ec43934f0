mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pipeline.py -m gpu -q -x -k "sweep_lanes or fused_pixel or graph_sweep" 2>&1 | tail -30 > gpurun_out/r2_t30.log
for L in 2 3 1; do
timeout 400 python bench.py --no-cpu-baseline --no-second-geometry --steps 200 --warmup 10 --in-flight $L > gpurun_out/r2_b30_L$L.json 2> gpurun_out/r2_b30_L$L.err
done
echo done
