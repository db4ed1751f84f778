export F=$PWD/tools/probe
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "maxpool" 2>&1 | tail -2
POPCORN_B200_LIB=$F/libpc_occ3.so timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "maxpool" 2>&1 | tail -2
for rep in 1 2; do for l in main ps0 nd16 occ3; do
  if [ $l = main ]; then unset POPCORN_B200_LIB; else export POPCORN_B200_LIB=$F/libpc_$l.so; fi
  echo -n "[$l] pool  "; KB_POOL=1 KB_ONLY=tc KB_ITERS=20 python tools/conv_layer_bench.py 8 8 4096 8192 | tail -1
  echo -n "[$l] pool16 "; KB_POOL=1 KB_ONLY=tc KB_ITERS=20 python tools/conv_layer_bench.py 16 16 2048 8192 | tail -1
done; done
