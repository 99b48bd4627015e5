#!/bin/bash
O=gpurun_out/${1:-san}
mkdir -p $O
for tool in memcheck synccheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool python tools/sanitize.py > $O/$tool.log 2>&1
  echo "$tool rc=$?" | tee -a $O/summary.txt
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY" $O/$tool.log | tee -a $O/summary.txt
done
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt; tail -3 $O/pytest_gpu.log | tee -a $O/summary.txt
python bench.py > $O/bench_default.json 2> $O/bench_default.err
python bench.py --impl reference --steps 40 --warmup 3 > $O/bench_reference.json 2> $O/bench_reference.err
python -c "
import json
d=json.load(open('$O/bench_default.json')); print('default value=%.4g e2e=%.4g cpu=%.4g (%s)' % (d['value'], d['e2e']['value'], d['cpu_baseline']['value'], d['cpu_baseline']['sample'][-60:]))
d=json.load(open('$O/bench_reference.json')); print('reference value=%.4g ms_per_step=%.4g' % (d['value'], d['ms_per_step']))" | tee -a $O/summary.txt
