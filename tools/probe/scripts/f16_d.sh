set -x
export F=$PWD/tools/probe
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for l in main occ3 occ5 occ9 win16 tf32 main; do
  if [ $l = main ]; then unset POPCORN_B200_LIB; else export POPCORN_B200_LIB=$F/libpc_$l.so; fi
  timeout 600 python bench.py --steps 3 --warmup 3 --skip-cpu-baseline --skip-timeseries --skip-e2e --skip-train --skip-ensemble --skip-gpu-baseline --skip-alone > gpurun_out/f16d_bench_$l.log 2>&1; tail -c 200 gpurun_out/f16d_bench_$l.log; echo
done
