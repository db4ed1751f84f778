set -x
export F=$PWD/tools/probe
POPCORN_B200_LIB=$F/libpc_x2.so timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
for l in main x2 main x2; do
  if [ $l = main ]; then unset POPCORN_B200_LIB; else export POPCORN_B200_LIB=$F/libpc_$l.so; fi
  echo "[$l]"; python tools/head_bench.py 2>&1 | tail -1; KB_ONLY=tc KB_ITERS=20 python tools/conv_layer_bench.py 8 8 4096 8192 | tail -1; KB_ONLY=tc KB_ITERS=20 python tools/conv_layer_bench.py 16 16 2048 8192 | tail -1
done
for l in main x2; do
  if [ $l = main ]; then unset POPCORN_B200_LIB; else export POPCORN_B200_LIB=$F/libpc_$l.so; fi
  timeout 600 python bench.py --steps 3 --warmup 3 --skip-cpu-baseline --skip-timeseries --skip-e2e --skip-train --skip-ensemble --skip-gpu-baseline --skip-alone > gpurun_out/x2_bench_$l.log 2>&1; tail -c 150 gpurun_out/x2_bench_$l.log; echo
done
