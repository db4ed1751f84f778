timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py tests/test_gpu_unet_train.py tests/test_gpu_country.py -m gpu -x -q 2>&1 | tail -3
for i in 1 2; do KB_ONLY=simt KB_ITERS=10 python tools/conv_layer_bench.py 2 8 4096 8192 | tail -1; KB_ONLY=simt KB_ITERS=10 python tools/conv_layer_bench.py 4 8 4096 8192 | tail -1; done
timeout 600 python bench.py --steps 5 --warmup 3 --skip-cpu-baseline --skip-timeseries --skip-e2e --skip-ensemble --skip-gpu-baseline --skip-alone > gpurun_out/mt_bench.log 2>&1; tail -c 150 gpurun_out/mt_bench.log
