#!/bin/bash
# Dev tool: time the step kernel of every library variant under build/variants (launch-bounds sweeps).
for lib in ${VARIANT_DIR:-build/variants}/*.so; do
  PVDER_B200_LIB=$PWD/$lib python bench.py --steps ${STEPS:-20} --warmup 3 --no-cpu-baseline --e2e-steps 1 ${BENCH_ARGS} 2>/dev/null | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lib', 'kernel_ms=%.4f' % d['roofline']['kernel_ms'], 'env-steps/s=%.4g' % d['value'], 'frac=%.3f' % d['roofline']['frac'])"
done
