#!/bin/bash
# Round-2 evidence set (one GPU): smoke, GPU tests, bench lines (both arms), launch list, ncu full + flop counters,
# 160-step trace, sanitizers.   usage: gpurun -- 'bash tools/r2_evidence.sh TAG'
O=gpurun_out/${1:-r2_evidence}
mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/summary.txt
timeout 900 python -m pytest tests -m gpu -q --timeout=300 --timeout-method=thread > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt; tail -3 $O/pytest_gpu.log | tee -a $O/summary.txt
python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?" | tee -a $O/summary.txt
python bench.py --steps 20 --warmup 5 > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?" | tee -a $O/summary.txt
python bench.py --steps 160 --warmup 5 --no-cpu-baseline --no-extra-configs > $O/bench_episode160.json 2> $O/bench_episode160.err
python bench.py --model model_2 --steps 40 --warmup 5 --no-cpu-baseline --no-extra-configs > $O/bench_model2_auto.json 2> $O/bench_model2_auto.err
python bench.py --model model_2 --three-phase-mode split --steps 40 --warmup 5 --no-cpu-baseline --no-extra-configs > $O/bench_model2_split.json 2> $O/bench_model2_split.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra-configs --e2e-steps 1 > $O/launches_bench.log 2>&1
M=smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_fp64_pred_on.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum
B="--steps 12 --warmup 3 --no-cpu-baseline --no-extra-configs --e2e-steps 1"
ncu --metrics $M --clock-control none -k "regex:^step_kernel$" -s 10 -c 1 --csv --log-file $O/flops_1ph.csv python bench.py $B > $O/f1.log 2>&1
ncu --metrics $M --clock-control none -k regex:step_kernel_split3 -s 10 -c 1 --csv --log-file $O/flops_split.csv python bench.py --model model_2 --three-phase-mode split $B > $O/f3.log 2>&1
ncu --metrics $M --clock-control none -k "regex:^step_kernel$" -s 10 -c 1 --csv --log-file $O/flops_m2auto.csv python bench.py --model model_2 $B > $O/f2.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:^step_kernel$" -s 10 -c 1 -o $O/step_1ph python bench.py $B > $O/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_kernel_split3 -s 10 -c 1 -o $O/step_split3 python bench.py --model model_2 --three-phase-mode split $B > $O/ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:^step_kernel$" -s 10 -c 1 -o $O/step_m2auto python bench.py --model model_2 $B > $O/ncu2.log 2>&1
python tools/step_trace.py model_1 > $O/step_trace_1ph.txt 2>&1
python tools/single_env_latency.py > $O/single_env_latency.json 2> $O/single.err
for tool in memcheck synccheck racecheck; do
  timeout 400 compute-sanitizer --tool $tool python tools/sanitize.py > $O/$tool.log 2>&1
  echo "$tool rc=$?" | tee -a $O/summary.txt
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY" $O/$tool.log | tee -a $O/summary.txt
done
for f in $O/bench_*.json; do python - $f <<'PY' | tee -a $O/summary.txt
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r = d.get('roofline') or {}; c = d.get('cpu_baseline') or {}
    print(sys.argv[1], 'value=%.4g ms/step=%.4g e2e=%.4g frac=%s yard=%s cpu=%s' % (d['value'], d['ms_per_step'], d['e2e']['value'], r.get('frac'), r.get('frac_yardstick'), c.get('value')))
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
done
ls -la $O | tee -a $O/summary.txt
