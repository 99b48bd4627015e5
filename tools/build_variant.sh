#!/bin/bash
# Dev tool: build a kernel variant into build/variants/NAME.so with extra nvcc flags, print spill/SASS stats.
# usage: tools/build_variant.sh NAME [extra nvcc flags...]      (KERNEL=regex of the mangled name to report)
NAME=$1; shift
K=${KERNEL:-step_kernelINS_8Model1phELb0}
cd /root/repo/gym-solarpvder-environment_b200/csrc
mkdir -p /root/repo/build/variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -diag-suppress 177 -shared -Xcompiler -fPIC -Xptxas -v "$@" -o /root/repo/build/variants/$NAME.so pvder_kernels.cu 2>&1 | grep -A2 "Function properties for _ZN5pvder[0-9]*$K" | cut -c1-160
cd /root/repo
python tools/sass_loops.py build/variants/$NAME.so $K 2>/dev/null | head -2 | tail -1
