#!/bin/bash
# Round-2 GPU call 1: same-box baseline of the round-1 tree, the fine-step builds at several register caps, GPU tests.
O=gpurun_out/r2a
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > $O/gpu.txt
run() {  # label, lib, dir, extra args...
  local label=$1 lib=$2 dir=$3; shift 3
  ( cd $dir && PVDER_B200_LIB=$lib python bench.py --steps 160 --warmup 5 --no-cpu-baseline --e2e-steps 1 "$@" 2>$OLDPWD/$O/$label.err ) > $O/$label.json
  python - "$label" $O/$label.json <<'PY' | tee -a $O/summary.txt
import sys, json
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print(sys.argv[1], 'kernel_ms=%.4f' % d['roofline']['kernel_ms'], 'env-steps/s=%.4g' % d['value'], 'windup=%s exact=%s' % (d['episode_stats']['windup_sub_steps'], d['episode_stats']['exact_sub_steps']))
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
}
R=$PWD
run r1_default $R/build/r1_tree/gym-solarpvder-environment_b200/csrc/libpvder_b200.so build/r1_tree
run r1_refine1 $R/build/variants/refine1.so build/r1_tree
for lib in build/variants/fine_e*.so; do
  n=$(basename $lib .so)
  run $n $R/$lib .
done
run fine_e_norefine $R/build/variants/fine_e.so . --cfg refine_input_level=0 --cfg startup_level=0
run fine_e_b96r224_norefine $R/build/variants/fine_e_b96r224.so . --cfg refine_input_level=0 --cfg startup_level=0
run fine_e_onaction $R/build/variants/fine_e.so . --cfg refine_on_action=true
run r1_default_again $R/build/r1_tree/gym-solarpvder-environment_b200/csrc/libpvder_b200.so build/r1_tree
run fine_e_model2 $R/build/variants/fine_e.so . --model model_2
run fine_e_split $R/build/variants/fine_e.so . --model model_2 --three-phase-mode split --steps 40
run r1_split $R/build/r1_tree/gym-solarpvder-environment_b200/csrc/libpvder_b200.so build/r1_tree --model model_2 --three-phase-mode split --steps 40
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
tail -30 $O/pytest_gpu.log | tee -a $O/summary.txt
