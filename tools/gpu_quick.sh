#!/bin/bash
# Dev tool: GPU tests + the three headline bench lines.
TAG=${1:-quick}
O=gpurun_out/$TAG
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
tail -15 $O/pytest_gpu.log | tee -a $O/summary.txt
python bench.py --steps 40 --no-cpu-baseline > $O/bench_model1.json 2> $O/bench_model1.err
python bench.py --model model_2 --steps 40 --no-cpu-baseline > $O/bench_model2_auto.json 2> $O/bench_model2_auto.err
python bench.py --model model_2 --three-phase-mode split --steps 40 --no-cpu-baseline > $O/bench_split.json 2> $O/bench_split.err
for f in $O/bench_*.json; do python -c "
import json,sys
d=json.load(open('$f')); print('$f', 'value=%.4g ms/step=%.4g e2e=%.4g frac=%.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac']))" | tee -a $O/summary.txt; done
