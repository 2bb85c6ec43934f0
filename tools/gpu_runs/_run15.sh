timeout 900 python -m pytest tests/test_gpu_sparse.py tests/test_gpu_dense.py tests/test_gpu_pipeline.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2_t15.log
for sk in 0 1 2; do
  P3D_F16_STREAMK=$sk timeout 300 python bench.py --no-cpu-baseline --no-second-geometry --steps 100 --warmup 10 > gpurun_out/r2_b15_sk$sk.json 2> gpurun_out/r2_b15_sk$sk.err
done
for fx in 3 10; do
  P3D_F16_SKFIX=$fx timeout 300 python bench.py --no-cpu-baseline --no-second-geometry --steps 100 --warmup 10 > gpurun_out/r2_b15_fix$fx.json 2> gpurun_out/r2_b15_fix$fx.err
done
echo done
