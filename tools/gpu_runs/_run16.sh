timeout 900 python -m pytest tests/test_gpu_dense.py tests/test_gpu_pipeline.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2_t16.log
timeout 300 python tools/dense_bench.py > gpurun_out/dense_bench_ws.jsonl 2>&1
P3D_DENSE_WS=0 timeout 300 python tools/dense_bench.py > gpurun_out/dense_bench_nows.jsonl 2>&1
timeout 300 python bench.py --no-cpu-baseline --no-second-geometry --steps 100 --warmup 10 > gpurun_out/r2_b16.json 2> gpurun_out/r2_b16.err
echo done
