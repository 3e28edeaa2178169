set -x
mkdir -p gpurun_out/r2
QLNLP_PDL=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2/pytest_al.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest_al.log
tail -3 gpurun_out/r2/pytest_al.log
timeout 1200 python tools/ab_bench.py run default default@QLNLP_PDL=1 default default@QLNLP_PDL=1 default default@QLNLP_PDL=1 > gpurun_out/r2/ab_al.log 2>&1
cat gpurun_out/r2/ab_al.log
