timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2_t9_all.log
timeout 400 python tools/op_bench.py > gpurun_out/r2_op_bench9.jsonl 2> gpurun_out/r2_op_bench9.err
P3D_BEV_POOL_VARIANT=1 timeout 200 python tools/op_bench.py --only bev_pool > gpurun_out/r2_op_bench9_bevold.jsonl 2> /dev/null
timeout 900 python bench.py --steps 100 > gpurun_out/r2_b9.json 2> gpurun_out/r2_b9.err
echo done
