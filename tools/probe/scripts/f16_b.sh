set -x
export F=$PWD/tools/probe/libpc_f16.so
POPCORN_B200_LIB=$F timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
for l in main f16 main f16; do
  if [ $l = main ]; then unset POPCORN_B200_LIB; else export POPCORN_B200_LIB=$F; fi
  echo "[$l]"; timeout 300 python tools/head_bench.py 2>&1 | tail -3
done
for l in main f16; do
  if [ $l = main ]; then unset POPCORN_B200_LIB; else export POPCORN_B200_LIB=$F; fi
  timeout 600 python bench.py --steps 3 --warmup 3 --skip-cpu-baseline --skip-timeseries --skip-e2e --skip-ensemble --skip-gpu-baseline --skip-alone > gpurun_out/f16b_bench_$l.log 2>&1; tail -c 300 gpurun_out/f16b_bench_$l.log; echo
done
