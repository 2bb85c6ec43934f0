mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2_t25_all.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke25.log 2>&1
timeout 300 python tools/op_bench.py --only voxelize > gpurun_out/r2_op25.jsonl 2>&1
echo done
