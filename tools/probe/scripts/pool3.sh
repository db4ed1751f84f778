export F=$PWD/tools/probe
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for l in ps0 main; do
  if [ $l = main ]; then unset POPCORN_B200_LIB; else export POPCORN_B200_LIB=$F/libpc_$l.so; fi
  timeout 600 python bench.py --steps 5 --warmup 3 --skip-cpu-baseline --skip-timeseries --skip-e2e --skip-train --skip-ensemble --skip-gpu-baseline --skip-alone > gpurun_out/pool3_bench_$l.log 2>&1; tail -c 150 gpurun_out/pool3_bench_$l.log; echo
done
