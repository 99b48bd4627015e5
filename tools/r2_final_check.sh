#!/bin/bash
# Round-2 final check of HEAD on one GPU: smoke, the whole GPU test tier, the two examples, the default bench line.
O=gpurun_out/${1:-r2_final}; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/summary.txt
timeout 900 python -m pytest tests -m gpu -q --timeout=300 --timeout-method=thread > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt; tail -2 $O/pytest_gpu.log | tee -a $O/summary.txt
timeout 300 python examples/random_agent.py > $O/example_random_agent.log 2>&1; echo "random_agent rc=$?" | tee -a $O/summary.txt; tail -3 $O/example_random_agent.log | tee -a $O/summary.txt
timeout 300 python examples/dqn_rollout.py > $O/example_dqn_rollout.log 2>&1; echo "dqn_rollout rc=$?" | tee -a $O/summary.txt; tail -3 $O/example_dqn_rollout.log | tee -a $O/summary.txt
python bench.py --steps 20 --warmup 5 > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?" | tee -a $O/summary.txt
python - $O/bench_default.json <<'PY' | tee -a $O/summary.txt
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); e = d['e2e']; r = d['roofline']
print('value=%.4g ms=%.4g e2e=%.4g frac=%.4g traffic=%s check=%s cpu=%.4g launches=%s' % (d['value'], d['ms_per_step'], e['value'], r['frac'], r['traffic'], r['profile_check'], d['cpu_baseline']['value'], d['gpu_launches']))
PY
