timeout 300 python -m pytest tests/test_gpu_boxes.py tests/test_gpu_pipeline.py -m gpu -q 2>&1 | tail -6 > gpurun_out/r2_t10.log
timeout 300 python tools/op_bench.py --only bev_pool,c4,c2 > gpurun_out/r2_op_bench10.jsonl 2> gpurun_out/r2_op_bench10.err
P3D_DENSE_NTILE=64 timeout 300 python tools/dense_bench.py > gpurun_out/r2_dense_bench10_n64.jsonl 2> gpurun_out/r2_dense_bench10_n64.err
P3D_DENSE_NTILE=64 P3D_DENSE_MT=1 timeout 300 python tools/dense_bench.py > gpurun_out/r2_dense_bench10_n64_mt1.jsonl 2> /dev/null
timeout 600 python bench.py --steps 100 --no-second-geometry > gpurun_out/r2_b10.json 2> gpurun_out/r2_b10.err
echo done
