for l in a x1 x2 x16 x32 x48; do
  export POPCORN_B200_LIB=$PWD/tools/probe/libpc_tcprobe$l.so
  echo "== $l store"; KB_ONLY=tc KB_ITERS=10 python tools/conv_layer_bench.py 8 8 4096 8192 2>&1 | tail -4
  echo "== $l pool"; KB_POOL=1 KB_ONLY=tc KB_ITERS=10 python tools/conv_layer_bench.py 8 8 4096 8192 2>&1 | tail -4
done
unset POPCORN_B200_LIB
echo "== main"; KB_ONLY=tc KB_ITERS=10 python tools/conv_layer_bench.py 8 8 4096 8192 2>&1 | tail -1; KB_POOL=1 KB_ONLY=tc KB_ITERS=10 python tools/conv_layer_bench.py 8 8 4096 8192 2>&1 | tail -1
