ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ncu_bench4.csv \
    python bench.py --steps 1 --warmup 3 --skip-cpu-baseline --skip-timeseries --skip-e2e --skip-train --skip-ensemble --skip-gpu-baseline --skip-alone > gpurun_out/ncu_bench4.log 2>&1
tail -c 200 gpurun_out/ncu_bench4.log
