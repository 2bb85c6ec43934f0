mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r2_t31_all.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke31.log 2>&1
timeout 900 python bench.py > gpurun_out/r2_b31.json 2> gpurun_out/r2_b31.err
timeout 600 python tools/op_bench.py > gpurun_out/r2_op31.jsonl 2>&1
timeout 300 ncu --set full --clock-control none -k regex:vox_ --launch-count 6 -o gpurun_out/r02_voxelize python tools/op_bench.py --only voxelize --c3-only --iters 1 > /dev/null 2> gpurun_out/r2_ncu31a.err
timeout 600 ncu --set full --clock-control none -k regex:'conv_f16_kernel|conv_wm_kernel' --launch-count 20 -o gpurun_out/r02_sparse python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-second-geometry --in-flight 1 > /dev/null 2> gpurun_out/r2_ncu31b.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-second-geometry --in-flight 1 > gpurun_out/r2_b31_ncu.log 2>&1
echo done
