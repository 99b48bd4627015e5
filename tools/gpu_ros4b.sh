mkdir -p gpurun_out/ros4b
STEPS=160 bash tools/variant_sweep.sh 2>&1 | tee gpurun_out/ros4b/sweep.txt
python bench.py --model model_2 --three-phase-mode split --steps 40 --no-cpu-baseline > gpurun_out/ros4b/bench_split.json 2> gpurun_out/ros4b/bench_split.err
python bench.py --model model_2 --steps 160 --no-cpu-baseline > gpurun_out/ros4b/bench_auto.json 2> gpurun_out/ros4b/bench_auto.err
for f in gpurun_out/ros4b/bench_*.json; do python -c "
import json,sys
d=json.load(open('$f')); print('$f', 'value=%.4g ms/step=%.4g e2e=%.4g frac=%.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac']))"; done
