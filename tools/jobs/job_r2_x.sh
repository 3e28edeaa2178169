set -x
mkdir -p gpurun_out/r2
QLNLP_LIB=$PWD/quadruped_landing_b200/libqlnlp_mra.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2/pytest_x.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest_x.log
tail -4 gpurun_out/r2/pytest_x.log
timeout 900 python tools/ab_bench.py run base default mra mrb nw12 default mra mrb > gpurun_out/r2/ab_x.log 2>&1
cat gpurun_out/r2/ab_x.log
