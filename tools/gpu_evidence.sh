#!/bin/bash
# Dev tool: the full single-GPU evidence set of a round (tests, bench lines, launch list, ncu captures,
# executed-flop counters, sanitizers, config 1 / config 5 runs).   usage: gpurun -- 'bash tools/gpu_evidence.sh TAG'
O=gpurun_out/${1:-evidence}
mkdir -p $O
bash tools/gpu_final.sh $1
M=smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_fp64_pred_on.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum
ncu --metrics $M --clock-control none -k regex:step_kernel -s 4 -c 1 --csv --log-file $O/flops_1ph.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $O/f1.log 2>&1
ncu --metrics $M --clock-control none -k regex:step_kernel_split3 -s 4 -c 1 --csv --log-file $O/flops_split.csv python bench.py --model model_2 --three-phase-mode split --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $O/f3.log 2>&1
python tools/single_env_latency.py > $O/single_env_latency.json 2> $O/single.err
python examples/dqn_rollout.py > $O/dqn_rollout.json 2> $O/dqn.err
for tool in memcheck synccheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool python tools/sanitize.py > $O/$tool.log 2>&1
  echo "$tool rc=$?" | tee -a $O/summary.txt
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY" $O/$tool.log | tee -a $O/summary.txt
done
cat $O/single_env_latency.json | tee -a $O/summary.txt
python -c "
import json
d=json.load(open('$O/dqn_rollout.json'))
for k,v in d.items():
    if isinstance(v, dict): print(k, 'env_steps_per_s=%.4g ms/it=%.4g' % (v['env_steps_per_s'], v['ms_per_iteration']))
    else: print(k, v)" | tee -a $O/summary.txt
