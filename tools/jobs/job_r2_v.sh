set -x
mkdir -p gpurun_out/r2
# parity of the new f-chain / ticket / flush code first (bit-exact f, fuzz), then A/B against the previous build
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2/pytest_v.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest_v.log
tail -4 gpurun_out/r2/pytest_v.log
timeout 900 python tools/ab_bench.py run base default nb1 nb3 base default > gpurun_out/r2/ab_v.log 2>&1
cat gpurun_out/r2/ab_v.log
