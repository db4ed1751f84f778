set -x
export F=$PWD/tools/probe/libpc_f16.so
POPCORN_B200_LIB=$F timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q 2>&1 | tail -15
POPCORN_B200_LIB=$F timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -15
bash tools/ab_layers.sh "main libpc_f16.so" "8 8 4096 8192" "8 16 4096 8192" "16 16 2048 8192" 2>&1 | grep -v "^+"
for l in main f16; do
  if [ $l = main ]; then unset POPCORN_B200_LIB; else export POPCORN_B200_LIB=$F; fi
  timeout 600 python bench.py --steps 3 --warmup 3 --skip-cpu-baseline --skip-timeseries --skip-e2e --skip-train --skip-ensemble --skip-gpu-baseline --skip-alone > gpurun_out/f16_bench_$l.log 2>&1; tail -c 300 gpurun_out/f16_bench_$l.log; echo
done
