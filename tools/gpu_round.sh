#!/bin/bash
# Dev tool: one gpurun call = GPU tests + bench lines + ncu launch list + ncu full captures.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_round.sh TAG'
TAG=${1:-rX}
O=gpurun_out/$TAG
mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/summary.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
tail -3 $O/pytest_gpu.log | tee -a $O/summary.txt
python bench.py > $O/bench_model1.json 2> $O/bench_model1.err; echo "bench1 rc=$?" | tee -a $O/summary.txt
python bench.py --model model_2 --no-cpu-baseline > $O/bench_model2_auto.json 2> $O/bench_model2_auto.err
python bench.py --model model_2 --three-phase-mode general --steps 40 --no-cpu-baseline > $O/bench_model2_general.json 2> $O/bench_model2_general.err
python bench.py --impl reference --steps 40 --warmup 3 > $O/bench_reference.json 2> $O/bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/launches.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $O/launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 4 -c 1 -o $O/step_1ph python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $O/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 4 -c 1 -o $O/step_3ph_general python bench.py --model model_2 --three-phase-mode general --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $O/ncu3.log 2>&1
ls -la $O | tee -a $O/summary.txt
