export F=$PWD/tools/probe
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
for i in 1 2; do python tools/head_bench.py 2>&1 | tail -1; done
for l in main mt42 main mt42; do
  if [ $l = main ]; then unset POPCORN_B200_LIB; else export POPCORN_B200_LIB=$F/libpc_$l.so; fi
  echo -n "[$l] "; KB_ONLY=simt KB_ITERS=10 python tools/conv_layer_bench.py 4 8 4096 8192 | tail -1
done
unset POPCORN_B200_LIB
timeout 600 python bench.py --steps 5 --warmup 3 --skip-cpu-baseline --skip-timeseries --skip-e2e --skip-train --skip-ensemble --skip-gpu-baseline --skip-alone > gpurun_out/head3_bench.log 2>&1; tail -c 150 gpurun_out/head3_bench.log
