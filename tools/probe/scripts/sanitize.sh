# compute-sanitizer over one census train step + one dense eval forward with census sums and a tile accumulation (tools/one_train_step.py)
for tool in racecheck synccheck memcheck; do
  timeout 900 compute-sanitizer --tool $tool python tools/one_train_step.py small > gpurun_out/san_$tool.log 2>&1
  echo "== $tool"; grep -E "SUMMARY|^ok|Error|error" gpurun_out/san_$tool.log | tail -4
done
timeout 900 compute-sanitizer --tool racecheck python tools/one_train_step.py small finetune > gpurun_out/san_racecheck_ft.log 2>&1
echo "== racecheck finetune"; grep -E "SUMMARY|^ok" gpurun_out/san_racecheck_ft.log | tail -3
