#!/bin/bash
# Round-2 call 20: soak run of all step kernels (tools/soak.py), bounded by its own timeout.
O=gpurun_out/${1:-r2_c20}; mkdir -p $O
timeout 400 python tools/soak.py > $O/soak.json 2> $O/soak.err; echo "soak rc=$?" | tee -a $O/summary.txt; cat $O/soak.json | tee -a $O/summary.txt; tail -3 $O/soak.err
