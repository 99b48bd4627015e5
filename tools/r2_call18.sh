#!/bin/bash
# Round-2 call 18: last micro-variant batch on the final kernel (A/B on one box): exp-only / rotation-only short side-input
# chains, lazy gains, clamp-free second instantiation, two-Newton reciprocal.
O=gpurun_out/${1:-r2_c18}; mkdir -p $O
run() { # lib tag args...
  lib=$1; tag=$2; shift 2
  PVDER_B200_LIB=$PWD/$lib timeout 180 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra-configs --e2e-steps 1 "$@" 2>/dev/null | \
    python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib $tag kernel_ms=%.4f value=%.4g' % (d['roofline']['kernel_ms'], d['value']))" | tee -a $O/summary.txt
}
for rep in 1 2; do
  for lib in build/variants/q_*.so; do run $lib model_1; done
done
for lib in build/variants/q_*.so; do run $lib model_1_160 --steps 160; done
