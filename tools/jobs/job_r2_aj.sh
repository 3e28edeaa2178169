set -x
mkdir -p gpurun_out/r2
timeout 1200 python tools/ab_bench.py run prev default dplain prev default dplain > gpurun_out/r2/ab_aj.log 2>&1
cat gpurun_out/r2/ab_aj.log
