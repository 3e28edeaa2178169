set -x
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest_z.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest_z.log
tail -4 gpurun_out/r2/pytest_z.log
timeout 1200 python tools/ab_bench.py run default base default base default base default > gpurun_out/r2/ab_z.log 2>&1
cat gpurun_out/r2/ab_z.log
