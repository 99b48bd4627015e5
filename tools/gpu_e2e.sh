# Dev tool: host-buffer pipeline check (chunked host API tests + default bench line + launch list)
O=gpurun_out/${1:-e2e}
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "chunked or host_handle" 2>&1 | tail -3
python bench.py --no-cpu-baseline > $O/bench_default.json 2> $O/bench_default.err
python bench.py --model model_2 --no-cpu-baseline > $O/bench_model2_auto.json 2> $O/bench_model2_auto.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 140 --csv --log-file $O/launches.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $O/launches_bench.log 2>&1
for f in $O/bench_*.json; do python -c "
import json
d=json.loads(open('$f').read().strip().splitlines()[-1]); r=d.get('roofline') or {}
print('$f', 'value=%.4g ms/step=%.4g e2e=%.4g kernel_ms_in_e2e=%.4g frac=%s executed=%s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['kernel_ms_in_e2e'], r.get('frac'), r.get('executed')))"; done
