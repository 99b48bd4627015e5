#!/bin/bash
# Round-2 call 10: FP64/ALU mix micro-benchmark, stage-4 delta A/B (single-phase + three-lane), GPU tests of the delta build.
O=gpurun_out/${1:-r2_c10}; mkdir -p $O
build/fp64_mix > $O/fp64_mix.txt 2>&1; cat $O/fp64_mix.txt
run() { # lib tag args...
  lib=$1; tag=$2; shift 2
  PVDER_B200_LIB=$PWD/$lib python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra-configs --e2e-steps 1 "$@" 2>/dev/null | \
    python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib $tag kernel_ms=%.4f value=%.4g' % (d['roofline']['kernel_ms'], d['value']))" | tee -a $O/summary.txt
}
for rep in 1 2; do
  for lib in build/variants/*.so; do run $lib model_1; done
done
for lib in build/variants/a_base.so build/variants/b_delta.so; do run $lib split --model model_2 --three-phase-mode split; run $lib m2auto --model model_2; done
run build/variants/b_delta.so model_1_160 --steps 160
run build/variants/a_base.so model_1_160 --steps 160
timeout 900 python -m pytest tests -m gpu -x -q --timeout=300 --timeout-method=thread > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt; tail -3 $O/pytest_gpu.log | tee -a $O/summary.txt
