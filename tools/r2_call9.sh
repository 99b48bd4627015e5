#!/bin/bash
O=gpurun_out/r2ac
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --timeout=200 --timeout-method=thread > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
tail -25 $O/pytest_gpu.log | cut -c1-250 | tee -a $O/summary.txt
python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extra-configs --e2e-steps 1 --model model_2 --three-phase-mode split > $O/bench_split.json 2> $O/bench_split.err
python -c "
import json
d=json.loads(open('$O/bench_split.json').read().strip().splitlines()[-1])
print('split kernel_ms=%.4f value=%.4g' % (d['roofline']['kernel_ms'], d['value']))" | tee -a $O/summary.txt
