timeout 300 python -m pytest tests/test_gpu_dense.py -m gpu -q -x 2>&1 | tail -40 > gpurun_out/r2_d7_a.log
timeout 300 python -m pytest tests/test_gpu_dense.py -m gpu -q 2>&1 | tail -30 > gpurun_out/r2_d7_b.log
timeout 300 python tools/dense_bench.py --variants --tf32 > gpurun_out/r2_dense_bench7.jsonl 2> gpurun_out/r2_dense_bench7.err
echo done
