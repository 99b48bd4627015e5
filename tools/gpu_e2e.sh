#!/bin/bash
O=gpurun_out/${1:-e2e}
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
tail -5 $O/pytest_gpu.log | tee -a $O/summary.txt
python bench.py --steps 80 --no-cpu-baseline > $O/bench_model1.json 2> $O/bench_model1.err
python bench.py --model model_2 --three-phase-mode split --steps 40 --no-cpu-baseline > $O/bench_split.json 2> $O/bench_split.err
ncu --set full --clock-control none --import-source on -k regex:step_kernel_split3 -s 4 -c 1 -o $O/step_split3 python bench.py --model model_2 --three-phase-mode split --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $O/ncu_split.log 2>&1
for f in $O/bench_*.json; do python -c "
import json
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', 'value=%.4g ms/step=%.4g e2e=%.4g kernel_ms_in_e2e=%.4g' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['kernel_ms_in_e2e']))" | tee -a $O/summary.txt; done
