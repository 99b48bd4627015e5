#!/bin/bash
# Round-2 scaling check on one 8-GPU box (gpurun --gpus 8): the driver's N = 8 launch of the default bench (weak scaling,
# 1 Mi envs per GPU; its line also carries config 4 as written = 1 Mi envs over the 8 ranks, and the e2e leg with the
# host-copy ceiling measured by all ranks at once), then N = 2.
O=gpurun_out/${1:-r2_scale8}; mkdir -p $O
for N in 8 2; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_n$N.json 2> $O/bench_n$N.err
  echo "N=$N rc=$?" | tee -a $O/summary.txt
done
for f in $O/bench_*.json; do python - $f <<'PY' | tee -a $O/summary.txt
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); e = d['e2e']; c = d.get('configs', {}).get('strong_config4', {})
    print(sys.argv[1], 'n_gpus=%d value=%.4g ms/step=%.4g e2e=%.4g chunks=%s q=%.3g ceiling_gbs=%.4g frac_of_min=%.3g compact=%.4g strong4=%.4g (%.4g ms)' % (
        d['n_gpus'], d['value'], d['ms_per_step'], e['value'], e.get('chunks'), e.get('copy_to_kernel_time_ratio', 0), e.get('host_copy_ceiling_gbs', 0),
        e.get('frac_of_min_device_rate_and_copy_ceiling', 0), (e.get('compact') or {}).get('value', 0), c.get('value', 0), c.get('ms_per_step', 0)))
except Exception as ex:
    print(sys.argv[1], 'FAILED', ex)
PY
done
