#!/bin/bash
# Round-2 call 12: Cramer 2x2 block pivot for the current pair on top of the new pivot order.
O=gpurun_out/${1:-r2_c12}; mkdir -p $O
run() { # lib tag args...
  lib=$1; tag=$2; shift 2
  PVDER_B200_LIB=$PWD/$lib python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra-configs --e2e-steps 1 "$@" 2>/dev/null | \
    python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib $tag kernel_ms=%.4f value=%.4g' % (d['roofline']['kernel_ms'], d['value']))" | tee -a $O/summary.txt
}
for rep in 1 2; do
  for lib in build/variants/f_nolazy_pivot.so build/variants/i_*.so build/variants/j_*.so; do [ -f $lib ] && run $lib model_1; done
done
for lib in build/variants/f_nolazy_pivot.so build/variants/i_*.so build/variants/j_*.so; do [ -f $lib ] && run $lib model_1_160 --steps 160; [ -f $lib ] && run $lib m2auto --model model_2; done
