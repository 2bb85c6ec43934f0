for cfg in "8 1 0" "8 2 0" "16 1 0" "4 1 0" "8 1 1" "8 1 2" "8 1 4"; do set -- $cfg; P3D_F16_NPW=$1 P3D_F16_NSUB=$2 P3D_F16_FLAGS=$3 timeout 200 python tools/f16_probe.py > gpurun_out/r2_p5_npw$1_nsub$2_f$3.jsonl 2> gpurun_out/r2_p5_npw$1_nsub$2_f$3.err; done
timeout 500 python -m pytest tests/test_gpu_sparse.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2_t5_sparse.log
P3D_F16_FLAGS=2 timeout 300 python -m pytest tests/test_gpu_sparse.py -m gpu -q -x -k "single_conv or resnet" 2>&1 | tail -8 > gpurun_out/r2_t5_sparse_f2.log
for cfg in "8 1 0" "8 2 0" "16 1 0" "8 1 2"; do set -- $cfg; P3D_F16_NPW=$1 P3D_F16_NSUB=$2 P3D_F16_FLAGS=$3 timeout 300 python bench.py --precision f16x3 --no-cpu-baseline --steps 100 > gpurun_out/r2_b5_npw$1_nsub$2_f$3.json 2> gpurun_out/r2_b5_npw$1_nsub$2_f$3.err; done
echo done
