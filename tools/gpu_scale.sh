#!/bin/bash
# Dev tool: scaling run on one box (gpurun --gpus 8): N = 1, 2, 4, 8, weak scaling, 1 Mi envs per GPU.
O=gpurun_out/${1:-scale}
mkdir -p $O
python bench.py --steps 80 --no-cpu-baseline > $O/bench_n1.json 2> $O/bench_n1.err
for N in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 80 --no-cpu-baseline > $O/bench_n$N.json 2> $O/bench_n$N.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 40 --model model_2 --three-phase-mode split --no-cpu-baseline > $O/bench_split_n8.json 2> $O/bench_split_n8.err
for f in $O/bench_*.json; do python -c "
import json
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', 'n_gpus=%d value=%.4g ms/step=%.4g e2e=%.4g' % (d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value']))" | tee -a $O/summary.txt; done
