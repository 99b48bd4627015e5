#!/bin/bash
O=gpurun_out/${1:-sweeps}
mkdir -p $O
STEPS=20 BENCH_ARGS="--model model_2 --three-phase-mode split" bash tools/variant_sweep.sh 2>&1 | tee $O/sweep_split.txt
python bench.py --steps 80 --no-cpu-baseline > $O/bench_model1.json 2> $O/bench_model1.err
python -c "
import json
d=json.load(open('$O/bench_model1.json')); print('model1 value=%.4g e2e=%.4g kernel_ms_in_e2e=%.4g' % (d['value'], d['e2e']['value'], d['e2e']['kernel_ms_in_e2e']))" | tee -a $O/summary.txt
