#!/bin/bash
# Round-2 GPU call 2: structural variants of the segment loop, n_sim=60 slope runs, ncu of r1 vs new hot loop.
O=gpurun_out/r2b
mkdir -p $O
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
R1=$R/build/r1_tree/gym-solarpvder-environment_b200/csrc/libpvder_b200.so
run r1_default $R1 build/r1_tree
for v in e2 f g h; do
  run fine_$v $R/build/variants/fine_$v.so .
  run fine_${v}_norefine $R/build/variants/fine_$v.so . --cfg refine_input_level=0 --cfg startup_level=0
done
run fine_f_onaction $R/build/variants/fine_f.so . --cfg refine_on_action=true
run fine_f_nostartup $R/build/variants/fine_f.so . --cfg startup_level=0
run r1_n60 $R1 build/r1_tree --n-sim 60 --steps 40
run fine_e2_n60_norefine $R/build/variants/fine_e2.so . --n-sim 60 --steps 40 --cfg refine_input_level=0 --cfg startup_level=0
run fine_f_n60_norefine $R/build/variants/fine_f.so . --n-sim 60 --steps 40 --cfg refine_input_level=0 --cfg startup_level=0
run r1_default_again $R1 build/r1_tree
# ncu: one launch each of the r1 kernel and the new one (no refinement: same work)
( cd build/r1_tree && PVDER_B200_LIB=$R1 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 8 -c 1 -o $R/$O/ncu_r1 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $R/$O/ncu_r1.log 2>&1 )
PVDER_B200_LIB=$R/build/variants/fine_e2.so ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 8 -c 1 -o $O/ncu_fine_e2 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --e2e-steps 1 --cfg refine_input_level=0 --cfg startup_level=0 > $O/ncu_fine_e2.log 2>&1
PVDER_B200_LIB=$R/build/variants/fine_f.so ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 8 -c 1 -o $O/ncu_fine_f python bench.py --steps 8 --warmup 3 --no-cpu-baseline --e2e-steps 1 --cfg refine_input_level=0 --cfg startup_level=0 > $O/ncu_fine_f.log 2>&1
ls -la $O
