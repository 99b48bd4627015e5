#!/bin/bash
# Round-2 GPU call 3: full GPU test tier on the new tree, default bench line (with the extra configs), split/auto timings.
O=gpurun_out/r2c
mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/summary.txt
timeout 1200 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
tail -40 $O/pytest_gpu.log | tee -a $O/summary.txt
python bench.py --steps 20 --warmup 5 > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?" | tee -a $O/summary.txt
python - $O/bench_default.json <<'PY' | tee -a $O/summary.txt
import sys, json
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print('value=%.4g ms/step=%.4f e2e=%.4g' % (d['value'], d['ms_per_step'], d['e2e']['value']), 'ceiling=%.4g GB/s frac=%.3f' % (d['e2e']['host_copy_ceiling_gbs'], d['e2e']['frac_of_min_device_rate_and_copy_ceiling']))
print('roofline', {k: d['roofline'][k] for k in ('achieved', 'frac', 'frac_yardstick', 'peak', 'profile_check')})
for k, v in d['configs'].items():
    print(k, {a: b for a, b in v.items() if a != 'workload'})
print('cpu', d['cpu_baseline'])
PY
python bench.py --steps 160 --warmup 5 --no-cpu-baseline --no-extra-configs --e2e-steps 1 > $O/bench_160.json 2> $O/bench_160.err
python bench.py --steps 160 --warmup 5 --no-cpu-baseline --no-extra-configs --e2e-steps 1 --model model_2 > $O/bench_160_m2.json 2> $O/bench_160_m2.err
python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extra-configs --e2e-steps 1 --model model_2 --three-phase-mode split > $O/bench_split.json 2> $O/bench_split.err
for f in bench_160 bench_160_m2 bench_split; do python -c "
import json
d=json.loads(open('$O/$f.json').read().strip().splitlines()[-1])
print('$f', 'kernel_ms=%.4f value=%.4g' % (d['roofline']['kernel_ms'], d['value']), d['episode_stats'])" | tee -a $O/summary.txt; done
