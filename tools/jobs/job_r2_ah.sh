set -x
mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "known_answers or c2_batch or other_horizons or ragged or stress" > gpurun_out/r2/pytest_ah.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest_ah.log
tail -3 gpurun_out/r2/pytest_ah.log
timeout 1200 python tools/ab_bench.py run plain default plain default plain default > gpurun_out/r2/ab_ah.log 2>&1
cat gpurun_out/r2/ab_ah.log
