#!/bin/bash
# compute-sanitizer over the hand-written kernels at small shapes (SURVEY.md section 5 "race detection / sanitizers").
# usage (on a GPU box): bash profiles/scripts/sanitize.sh [outdir]   -> <outdir>/sanitize_{memcheck,racecheck,synccheck}.log
OUT=${1:-gpurun_out}
mkdir -p "$OUT"
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 0 \
      python profiles/scripts/sanitize_smoke.py > "$OUT/sanitize_$tool.log" 2>&1
  echo "== $tool: exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_smoke done|Error|hazard" "$OUT/sanitize_$tool.log" | sort | uniq -c | head -12
done
