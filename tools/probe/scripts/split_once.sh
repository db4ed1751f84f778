set -x
export F=$PWD/tools/probe
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py tests/test_gpu_country.py -m gpu -x -q 2>&1 | tail -5
bash tools/ab_layers.sh "main libpc_nosplit1.so" "8 8 4096 8192" "8 16 4096 8192" "16 16 2048 8192" 2>&1 | grep -v "^+"
for l in nosplit1 main; do
  if [ $l = main ]; then unset POPCORN_B200_LIB; else export POPCORN_B200_LIB=$F/libpc_$l.so; fi
  timeout 600 python bench.py --steps 3 --warmup 3 --skip-cpu-baseline --skip-timeseries --skip-e2e --skip-train --skip-ensemble --skip-gpu-baseline --skip-alone > gpurun_out/so_bench_$l.log 2>&1; tail -c 200 gpurun_out/so_bench_$l.log; echo
done
