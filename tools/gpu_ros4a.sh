mkdir -p gpurun_out/ros4a
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/ros4a/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/ros4a/pytest_gpu.log
STEPS=160 bash tools/variant_sweep.sh 2>&1 | tee gpurun_out/ros4a/sweep.txt
