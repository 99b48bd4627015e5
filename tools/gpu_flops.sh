#!/bin/bash
O=gpurun_out/${1:-flops}
mkdir -p $O
M=smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_fp64_pred_on.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum
ncu --metrics $M --clock-control none -k regex:step_kernel -s 4 -c 1 --csv --log-file $O/flops_1ph.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $O/f1.log 2>&1
ncu --metrics $M --clock-control none -k regex:step_kernel_split3 -s 4 -c 1 --csv --log-file $O/flops_split.csv python bench.py --model model_2 --three-phase-mode split --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $O/f3.log 2>&1
python tools/single_env_latency.py > $O/single_env_latency.json 2> $O/single.err
cat $O/single_env_latency.json
tail -8 $O/flops_1ph.csv; tail -8 $O/flops_split.csv
