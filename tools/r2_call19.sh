#!/bin/bash
# Round-2 call 19: compute-sanitizer on the extended tools/sanitize.py (partial-mask regions of the three-lane kernel, redo list).
O=gpurun_out/${1:-r2_c19}; mkdir -p $O
timeout 200 python tools/sanitize.py > $O/plain.log 2>&1; echo "plain rc=$?" | tee -a $O/summary.txt; tail -4 $O/plain.log | tee -a $O/summary.txt
for tool in synccheck racecheck memcheck; do
  timeout 500 compute-sanitizer --tool $tool python tools/sanitize.py > $O/$tool.log 2>&1
  echo "$tool rc=$?" | tee -a $O/summary.txt
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY" $O/$tool.log | tee -a $O/summary.txt
done
