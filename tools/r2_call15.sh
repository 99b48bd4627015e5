#!/bin/bash
# Round-2 call 15: three-lane kernel with the phase sums through shared memory (batched) vs warp shuffles, A/B + its GPU tests.
O=gpurun_out/${1:-r2_c15}; mkdir -p $O
run() { # lib tag args...
  lib=$1; tag=$2; shift 2
  PVDER_B200_LIB=$PWD/$lib timeout 180 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra-configs --e2e-steps 1 "$@" 2>/dev/null | \
    python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib $tag kernel_ms=%.4f value=%.4g' % (d['roofline']['kernel_ms'], d['value']))" | tee -a $O/summary.txt
}
for rep in 1 2; do
  for lib in build/variants/m_*.so; do run $lib split --model model_2 --three-phase-mode split; done
done
for lib in build/variants/m_*.so; do run $lib split_unbal --model model_2 --three-phase-mode split --grid-unbalance 0.95 1.03; run $lib m2auto --model model_2; done
PVDER_B200_LIB=$PWD/build/variants/m_smemsum.so timeout 600 python -m pytest tests -m gpu -x -q --timeout=200 --timeout-method=thread -k "split or unbalanced or three_phase or quarantin or redo or auto or modes or trajectory or smoke or golden or emulation" > $O/pytest_split.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt; tail -3 $O/pytest_split.log | tee -a $O/summary.txt
