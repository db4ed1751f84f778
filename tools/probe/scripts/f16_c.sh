set -x
export F=$PWD/tools/probe
POPCORN_B200_LIB=$F/libpc_f16n3.so timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5
for l in main f16 f16n3 main f16 f16n3; do
  if [ $l = main ]; then unset POPCORN_B200_LIB; else export POPCORN_B200_LIB=$F/libpc_$l.so; fi
  echo "[$l]"; timeout 300 python tools/head_bench.py 2>&1 | tail -1
done
