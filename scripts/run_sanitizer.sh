#!/bin/bash
# compute-sanitizer over every kernel of libtrt_b200 (scripts/sanitize_small.py); keeps the summaries for profiles/r02_sanitizer.txt
# usage (GPU box): bash scripts/run_sanitizer.sh gpurun_out/r02_sanitizer.txt
out=${1:-gpurun_out/r02_sanitizer.txt}
: > "$out"
for tool in memcheck racecheck initcheck synccheck; do
  echo "==== compute-sanitizer --tool $tool python scripts/sanitize_small.py" >> "$out"
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_small.py > /tmp/san_$tool.log 2>&1
  echo "exit code $?" >> "$out"
  grep -E "^(demo|200|2\+3|fused|flags|orbit|probes)" /tmp/san_$tool.log >> "$out"
  grep -E "=========" /tmp/san_$tool.log | grep -vE "^========= *$" | head -60 >> "$out"
done
tail -5 "$out"
