set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for i in 1 2; do timeout 300 python tools/head_bench.py 2>&1 | tail -1; done
