mkdir -p gpurun_out
timeout 300 python tools/wm_determinism.py > gpurun_out/r2_wm_det.log 2>&1
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_sparse.py -m gpu -q -x -k "narrow and True" 2>&1 | tail -40 > gpurun_out/r2_race24.log
timeout 600 compute-sanitizer --tool initcheck --print-limit 20 python -m pytest tests/test_gpu_sparse.py -m gpu -q -x -k "narrow and True" 2>&1 | tail -40 > gpurun_out/r2_init24.log
echo done
