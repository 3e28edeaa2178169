set -x
mkdir -p gpurun_out/r2
timeout 900 python tools/ab_bench.py run default default@QLNLP_BLOCKS_PER_SM=5 default@QLNLP_BLOCKS_PER_SM=7 default@QLNLP_BLOCKS_PER_SM=8 default > gpurun_out/r2/ab_ba.log 2>&1
cat gpurun_out/r2/ab_ba.log
