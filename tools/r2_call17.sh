#!/bin/bash
# Round-2 call 17: limit flags of the clamp carried across sub-steps (evaluated at the end of the step) vs at the loop head, A/B on one box + GPU tests of the new build.
O=gpurun_out/${1:-r2_c17}; mkdir -p $O
run() { # lib tag args...
  lib=$1; tag=$2; shift 2
  PVDER_B200_LIB=$PWD/$lib timeout 180 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra-configs --e2e-steps 1 "$@" 2>/dev/null | \
    python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib $tag kernel_ms=%.4f value=%.4g' % (d['roofline']['kernel_ms'], d['value']))" | tee -a $O/summary.txt
}
for rep in 1 2; do
  for lib in build/variants/o_*.so; do run $lib model_1; done
done
for lib in build/variants/o_*.so; do run $lib model_1_160 --steps 160; run $lib m2auto --model model_2; done
PVDER_B200_LIB=$PWD/build/variants/o_carry.so timeout 900 python -m pytest tests -m gpu -x -q --timeout=300 --timeout-method=thread > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt; tail -3 $O/pytest_gpu.log | tee -a $O/summary.txt
