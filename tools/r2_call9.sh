#!/bin/bash
O=gpurun_out/r2ab
mkdir -p $O
for lib in default r1commit; do
  if [ $lib = default ]; then L=$PWD/gym-solarpvder-environment_b200/csrc/libpvder_b200.so; else L=$PWD/build/variants/fix_r1commit.so; fi
  PVDER_B200_LIB=$L timeout 60 python tools/r2_debug_s1.py > $O/s1_$lib.log 2>&1; echo "$lib S1 rc=$?" | tee -a $O/summary.txt
  PVDER_B200_LIB=$L timeout 60 python tools/r2_debug_mild.py > $O/mild_$lib.log 2>&1; echo "$lib mild rc=$?" | tee -a $O/summary.txt
  PVDER_B200_LIB=$L timeout 60 python tools/r2_debug_hang.py split oddclamp 1000 > $O/clamp_$lib.log 2>&1; echo "$lib blow-up split rc=$?" | tee -a $O/summary.txt
  PVDER_B200_LIB=$L timeout 60 python tools/r2_debug_hang.py auto oddclamp 1000 > $O/clampa_$lib.log 2>&1; echo "$lib blow-up auto rc=$?" | tee -a $O/summary.txt
  PVDER_B200_LIB=$L python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extra-configs --e2e-steps 1 --model model_2 --three-phase-mode split > $O/bench_$lib.json 2> $O/bench_$lib.err
  python -c "
import json
d=json.loads(open('$O/bench_$lib.json').read().strip().splitlines()[-1])
print('$lib split kernel_ms=%.4f value=%.4g' % (d['roofline']['kernel_ms'], d['value']))" | tee -a $O/summary.txt
done
timeout 900 python -m pytest tests -m gpu -q --timeout=200 --timeout-method=thread > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
tail -15 $O/pytest_gpu.log | cut -c1-250 | tee -a $O/summary.txt
