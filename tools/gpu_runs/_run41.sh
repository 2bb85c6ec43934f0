mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke41.log 2>&1
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > gpurun_out/r2_t41_all.log
echo done
