set -x
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py tests/test_gpu_unet_train.py -m gpu -x -q 2>&1 | tail -4
for i in 1 2; do KB_ONLY=simt KB_ITERS=10 python tools/conv_layer_bench.py 2 8 4096 8192 | tail -1; KB_ONLY=simt KB_ITERS=10 python tools/conv_layer_bench.py 4 8 4096 8192 | tail -1; done
timeout 600 python bench.py --steps 3 --warmup 3 --skip-cpu-baseline --skip-timeseries --skip-e2e --skip-ensemble --skip-gpu-baseline --skip-alone > gpurun_out/lds_bench.log 2>&1; tail -c 200 gpurun_out/lds_bench.log
