set -x
mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sparse_true or host or known or stress or c2_batch" > gpurun_out/r2/pytest_bc.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest_bc.log
tail -2 gpurun_out/r2/pytest_bc.log
timeout 900 python tools/ab_bench.py run noil default noil default > gpurun_out/r2/ab_bc.log 2>&1
cat gpurun_out/r2/ab_bc.log
