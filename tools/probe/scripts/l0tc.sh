timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py tests/test_gpu_country.py -m gpu -x -q 2>&1 | tail -4
for v in 0 1; do
  POPCORN_CONV_TC_L0=$v timeout 600 python bench.py --steps 5 --warmup 3 --skip-cpu-baseline --skip-timeseries --skip-e2e --skip-train --skip-ensemble --skip-gpu-baseline --skip-alone > gpurun_out/l0tc_bench_$v.log 2>&1; tail -c 150 gpurun_out/l0tc_bench_$v.log; echo
done
