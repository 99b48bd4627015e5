#!/bin/bash
# Round-2 call 14: compaction kernel on the high-priority copy stream -- host-API GPU tests + the default bench line.
O=gpurun_out/${1:-r2_c14}; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q -k "host or compact or pipeline or handle" --timeout=300 > $O/pytest_host.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt; tail -3 $O/pytest_host.log | tee -a $O/summary.txt
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra-configs > $O/bench_default.json 2> $O/bench_default.err
python - $O/bench_default.json <<'PY' | tee -a $O/summary.txt
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); e = d['e2e']
print('value=%.4g ms=%.4g e2e=%.4g compact=%.4g frac=%s check=%s' % (d['value'], d['ms_per_step'], e['value'], e['compact']['value'], d['roofline']['frac'], d['roofline']['profile_check']))
PY
