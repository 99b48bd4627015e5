#!/bin/bash
# Round-2 call 11: single-phase variants on top of the stage-4 delta (eager gains, hot loop unrolled by 2, Vdc/dl pivoted before
# the currents, pre-scaled U rows, 64-thread CTAs).
O=gpurun_out/${1:-r2_c11}; mkdir -p $O
run() { # lib tag args...
  lib=$1; tag=$2; shift 2
  PVDER_B200_LIB=$PWD/$lib python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra-configs --e2e-steps 1 "$@" 2>/dev/null | \
    python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib $tag kernel_ms=%.4f value=%.4g' % (d['roofline']['kernel_ms'], d['value']))" | tee -a $O/summary.txt
}
for rep in 1 2; do
  for lib in build/variants/*.so; do run $lib model_1; done
done
for lib in build/variants/*.so; do run $lib model_1_160 --steps 160; done
for lib in build/variants/c_nolazy.so build/variants/f_nolazy_pivot.so build/variants/g_nolazy_pivot_su.so; do run $lib m2auto --model model_2; done
