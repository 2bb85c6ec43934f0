timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r2_t8_all.log
timeout 300 python tools/dense_bench.py > gpurun_out/r2_dense_bench8.jsonl 2> gpurun_out/r2_dense_bench8.err
timeout 600 python bench.py --steps 100 > gpurun_out/r2_b8.json 2> gpurun_out/r2_b8.err
echo done
