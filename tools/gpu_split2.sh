#!/bin/bash
O=gpurun_out/${1:-split2}
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "split or unbalanced or emulation or trajectory" > $O/pytest_split.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt; tail -3 $O/pytest_split.log | tee -a $O/summary.txt
python bench.py --model model_2 --three-phase-mode split --steps 40 --no-cpu-baseline > $O/bench_split.json 2> $O/bench_split.err
python -c "
import json
d=json.load(open('$O/bench_split.json')); print('split value=%.4g ms/step=%.4g frac=%.3f' % (d['value'], d['ms_per_step'], d['roofline']['frac']))" | tee -a $O/summary.txt
