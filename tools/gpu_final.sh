#!/bin/bash
# Dev tool: round evidence -- all GPU tests, smoke, the default bench lines, reference arm, launch list, ncu full.
# FAST=1: tests, smoke, default bench line and launch list only.
O=gpurun_out/${1:-final}
mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/summary.txt
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt; tail -3 $O/pytest_gpu.log | tee -a $O/summary.txt
python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?" | tee -a $O/summary.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file $O/launches.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $O/launches_bench.log 2>&1
if [ -z "$FAST" ]; then
python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?" | tee -a $O/summary.txt
python bench.py --model model_2 --no-cpu-baseline > $O/bench_model2_auto.json 2> $O/bench_model2_auto.err
python bench.py --model model_2 --three-phase-mode split --no-cpu-baseline > $O/bench_model2_split.json 2> $O/bench_model2_split.err
python bench.py --model model_2 --three-phase-mode split --grid-unbalance 0.95 1.03 --steps 40 --no-cpu-baseline > $O/bench_model2_split_unbal.json 2> $O/bench_model2_split_unbal.err
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 4 -c 1 -o $O/step_1ph python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $O/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_kernel_split3 -s 4 -c 1 -o $O/step_split3 python bench.py --model model_2 --three-phase-mode split --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $O/ncu3.log 2>&1
fi
for f in $O/bench_*.json; do python -c "
import json
d=json.loads(open('$f').read().strip().splitlines()[-1]); r=d.get('roofline') or {}; c=d.get('cpu_baseline') or {}
print('$f', 'value=%.4g ms/step=%.4g e2e=%.4g frac=%s cpu=%s' % (d['value'], d['ms_per_step'], d['e2e']['value'], r.get('frac'), c.get('value')))" | tee -a $O/summary.txt; done
