set -x
mkdir -p gpurun_out/r2
tools/micro/dfma > gpurun_out/r2/dfma_w.log 2>&1; cat gpurun_out/r2/dfma_w.log
# parity of the bulk-flush variant (bit-exact g / grad against the oracle), then A/B
QLNLP_LIB=$PWD/quadruped_landing_b200/libqlnlp_bf2.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2/pytest_w.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest_w.log
tail -4 gpurun_out/r2/pytest_w.log
timeout 900 python tools/ab_bench.py run base default bf1 bf2 default bf1 bf2 > gpurun_out/r2/ab_w.log 2>&1
cat gpurun_out/r2/ab_w.log
