#!/bin/bash
# Round-2 bring-up (after run_pair_probe.sh has confirmed the descriptor semantics):
#   gpurun --timeout 600 -- 'bash tools/probe/run_conv_pair.sh > gpurun_out/conv_pair.txt 2>&1'
cd "$(dirname "$0")/../.."
bash tools/probe/build_conv_pair.sh 2>&1 | grep -i "error" && exit 1
python tools/probe/emulate_conv_pair.py
echo "=== conv_ss (3xTF32, SS form) ==="
timeout 180 python tools/probe/test_conv_pair.py ss
echo "exit code $?"
for cp in 0 1; do
  echo "=== POPCORN_PAIR_A_CP=$cp ==="
  POPCORN_PAIR_A_CP=$cp timeout 180 python tools/probe/test_conv_pair.py
  echo "exit code $?"
done
echo "=== DDA feature pass, chunk layout vs shipped (A/B) ==="
timeout 300 python tools/probe/test_dda_c4.py 2048 2048
echo "exit code $?"
