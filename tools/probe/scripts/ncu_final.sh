set -x
# (1) DRAM bytes + durations of every launch of one bench step (bench.py's own launches: roofline.traffic comes from this capture)
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ncu_bench3.csv \
    python bench.py --steps 1 --warmup 3 --skip-cpu-baseline --skip-timeseries --skip-e2e --skip-train --skip-ensemble --skip-gpu-baseline --skip-alone > gpurun_out/ncu_bench3.log 2>&1
tail -c 200 gpurun_out/ncu_bench3.log
# (2) ncu --set full of the head (fp16 halves, three contexts) and of the 8->8 conv layer, store and pool epilogues
ncu --set full --import-source on --clock-control none -k regex:head_tc -s 2 -c 1 -o gpurun_out/r2_head_f16 -f python tools/head_bench.py > gpurun_out/ncu_head_f16.log 2>&1
KB_ONLY=tc KB_ITERS=2 ncu --set full --import-source on --clock-control none -k regex:conv3x3_tc -s 2 -c 1 -o gpurun_out/r2_conv_f16_store -f python tools/conv_layer_bench.py 8 8 4096 8192 > gpurun_out/ncu_conv_store.log 2>&1
KB_POOL=1 KB_ONLY=tc KB_ITERS=2 ncu --set full --import-source on --clock-control none -k regex:conv3x3_tc -s 2 -c 1 -o gpurun_out/r2_conv_f16_pool -f python tools/conv_layer_bench.py 8 8 4096 8192 > gpurun_out/ncu_conv_pool.log 2>&1
ls -la gpurun_out/*.ncu-rep
