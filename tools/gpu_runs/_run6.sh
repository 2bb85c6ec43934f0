timeout 300 python -m pytest tests/test_gpu_dense.py -m gpu -q -k f16_vs_oracle 2>&1 | tail -15 > gpurun_out/r2_d6_a.log
P3D_DENSE_BO=0 timeout 300 python -m pytest tests/test_gpu_dense.py -m gpu -q -k f16_vs_oracle 2>&1 | tail -15 > gpurun_out/r2_d6_b.log
P3D_DENSE_PITCH=16 timeout 300 python -m pytest tests/test_gpu_dense.py -m gpu -q -k f16_vs_oracle 2>&1 | tail -15 > gpurun_out/r2_d6_c.log
P3D_DENSE_PITCH=16 P3D_DENSE_BO=0 timeout 300 python -m pytest tests/test_gpu_dense.py -m gpu -q -k f16_vs_oracle 2>&1 | tail -15 > gpurun_out/r2_d6_d.log
timeout 300 python -m pytest tests/test_gpu_dense.py -m gpu -q 2>&1 | tail -15 > gpurun_out/r2_d6_e.log
timeout 300 python tools/dense_bench.py --variants --tf32 > gpurun_out/r2_dense_bench6.jsonl 2> gpurun_out/r2_dense_bench6.err
P3D_DENSE_PITCH=16 timeout 300 python tools/dense_bench.py > gpurun_out/r2_dense_bench6_p16.jsonl 2> gpurun_out/r2_dense_bench6_p16.err
echo done
