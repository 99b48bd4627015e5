#!/bin/bash
# Round-2 scaling check at N = 4 (gpurun --gpus 4): the driver's launch of the default bench.
O=gpurun_out/${1:-r2_scale4}; mkdir -p $O
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_n4.json 2> $O/bench_n4.err
echo "N=4 rc=$?" | tee -a $O/summary.txt
python - $O/bench_n4.json <<'PY' | tee -a $O/summary.txt
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); e = d['e2e']; c = d.get('configs', {}).get('strong_config4', {})
print('n_gpus=%d value=%.4g ms/step=%.4g e2e=%.4g chunks=%s q=%.3g ceiling_gbs=%.4g frac_of_min=%.3g compact=%.4g strong4=%.4g (%.4g ms)' % (
    d['n_gpus'], d['value'], d['ms_per_step'], e['value'], e.get('chunks'), e.get('copy_to_kernel_time_ratio', 0), e.get('host_copy_ceiling_gbs', 0),
    e.get('frac_of_min_device_rate_and_copy_ceiling', 0), (e.get('compact') or {}).get('value', 0), c.get('value', 0), c.get('ms_per_step', 0)))
PY
