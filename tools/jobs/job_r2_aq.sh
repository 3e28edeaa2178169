set -x
mkdir -p gpurun_out/r2
timeout 1200 python tools/ab_bench.py run default default@QLNLP_GRID=1024 default@QLNLP_GRID=820 default@QLNLP_GRID=1100 default > gpurun_out/r2/ab_aq.log 2>&1
cat gpurun_out/r2/ab_aq.log
