#!/bin/bash
O=gpurun_out/${1:-policy}
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt; tail -15 $O/pytest_gpu.log | tee -a $O/summary.txt
python examples/dqn_rollout.py > $O/dqn_rollout.json 2> $O/dqn.err; tail -3 $O/dqn.err
python -c "
import json
d=json.load(open('$O/dqn_rollout.json'))
for k,v in d.items():
    if isinstance(v, dict): print(k, 'env_steps_per_s=%.4g ms/it=%.4g' % (v['env_steps_per_s'], v['ms_per_iteration']))
    else: print(k, v)" | tee -a $O/summary.txt
