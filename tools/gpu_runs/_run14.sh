timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/r2_t14_all.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke14.log 2>&1
timeout 900 python bench.py > gpurun_out/r2_b14.json 2> gpurun_out/r2_b14.err
timeout 400 ncu --set full --clock-control none -k regex:'vox_|scat_|bev_fwd|prep_' --launch-count 24 -o gpurun_out/r02_hbm_ops python tools/op_bench.py --iters 1 --only voxelize,scatter,bev_pool > /dev/null 2> gpurun_out/r2_ncu14.err
echo done
