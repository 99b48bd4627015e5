#!/bin/bash
# Round-2 GPU call 4: full GPU test tier (all failures), bench default line.
O=gpurun_out/r2d
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
tail -60 $O/pytest_gpu.log | tee -a $O/summary.txt
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?" | tee -a $O/summary.txt
python - $O/bench_default.json <<'PY' | tee -a $O/summary.txt
import sys, json
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print('value=%.4g ms/step=%.4f e2e=%.4g compact=%.4g' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['compact']['value']), 'ceiling=%.4g GB/s frac=%.3f chunks=%d ratio=%.2f' % (d['e2e']['host_copy_ceiling_gbs'], d['e2e']['frac_of_min_device_rate_and_copy_ceiling'], d['e2e']['chunks'], d['e2e']['copy_to_kernel_time_ratio']))
for k, v in d['configs'].items():
    print(k, {a: b for a, b in v.items() if a != 'workload'})
PY
