#!/bin/bash
# Dev tool: GPU check of the three-lane (split) three-phase kernel: tests, bench lines, ncu full.
TAG=${1:-split}
O=gpurun_out/$TAG
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
tail -15 $O/pytest_gpu.log | tee -a $O/summary.txt
python bench.py --model model_2 --three-phase-mode split --steps 40 --no-cpu-baseline > $O/bench_split.json 2> $O/bench_split.err; echo "bench split rc=$?" | tee -a $O/summary.txt
python bench.py --model model_2 --three-phase-mode split --grid-unbalance 0.95 1.03 --steps 40 --no-cpu-baseline > $O/bench_split_unbal.json 2> $O/bench_split_unbal.err
python bench.py --steps 40 --no-cpu-baseline > $O/bench_model1.json 2> $O/bench_model1.err
ncu --set full --clock-control none --import-source on -k regex:step_kernel_split3 -s 4 -c 1 -o $O/step_split3 python bench.py --model model_2 --three-phase-mode split --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $O/ncu_split.log 2>&1
for f in $O/bench_*.json; do python -c "
import json,sys
d=json.load(open('$f')); print('$f', 'value=%.4g ms/step=%.4g e2e=%.4g frac=%.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac']))" | tee -a $O/summary.txt; done
