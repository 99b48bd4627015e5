#!/bin/bash
# Dev tool: build a kernel variant into build/variants/NAME.so with extra nvcc flags, print 1-ph stats.
# usage: tools/build_variant.sh NAME [extra nvcc flags...]
NAME=$1; shift
cd /root/repo/gym-solarpvder-environment_b200/csrc
mkdir -p /root/repo/build/variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -diag-suppress 177 -shared -Xcompiler -fPIC -Xptxas -v "$@" -o /root/repo/build/variants/$NAME.so pvder_kernels.cu 2>&1 | grep -A2 "Function properties for _ZN5pvder11step_kernelINS_8Model1ph" | cut -c1-160
cd /root/repo
python tools/sass_loops.py build/variants/$NAME.so 'step_kernelINS_8Model1phELb0' | head -3
