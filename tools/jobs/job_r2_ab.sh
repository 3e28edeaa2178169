set -x
mkdir -p gpurun_out/r2
QLNLP_LIB=$PWD/quadruped_landing_b200/libqlnlp_both.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2/pytest_ab.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest_ab.log
tail -3 gpurun_out/r2/pytest_ab.log
timeout 1200 python tools/ab_bench.py run head default sth ldh both head default sth ldh both > gpurun_out/r2/ab_ab.log 2>&1
cat gpurun_out/r2/ab_ab.log
