set -x
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2/pytest_bb.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest_bb.log
tail -2 gpurun_out/r2/pytest_bb.log
timeout 900 python tools/ab_bench.py run noil default noil default noil default > gpurun_out/r2/ab_bb.log 2>&1
cat gpurun_out/r2/ab_bb.log
