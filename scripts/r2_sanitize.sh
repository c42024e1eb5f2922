#!/bin/bash
# compute-sanitizer over every kernel on small problems (run under gpurun); output -> gpurun_out/r2_sanitizer.txt
out=gpurun_out/r2_sanitizer.txt
echo "== memcheck" > $out
timeout 1500 compute-sanitizer --tool memcheck python scripts/sanitize_smoke.py 2>&1 | grep -E "^ok|ERROR SUMMARY|Invalid|Error" | tail -40 >> $out
echo "== racecheck" >> $out
timeout 1500 compute-sanitizer --tool racecheck python scripts/sanitize_smoke.py 2>&1 | grep -E "^ok|RACECHECK SUMMARY|hazard|Error" | tail -40 >> $out
cat $out | tail -50
