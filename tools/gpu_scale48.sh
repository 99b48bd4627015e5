#!/bin/bash
O=gpurun_out/${1:-scale48}
mkdir -p $O
for N in 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 80 --no-cpu-baseline > $O/bench_n$N.json 2> $O/bench_n$N.err
done
nvidia-smi topo -m > $O/topo.txt 2>&1; lscpu | head -30 > $O/lscpu.txt; numactl -H > $O/numa.txt 2>&1
for f in $O/bench_*.json; do python -c "
import json
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', 'n_gpus=%d value=%.4g ms/step=%.4g e2e=%.4g' % (d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value']), d['config'].get('host_affinity'))" | tee -a $O/summary.txt; done
