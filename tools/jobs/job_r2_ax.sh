set -x
mkdir -p gpurun_out/r2
timeout 1200 python tools/ab_bench.py run default pfall default pfall default pfall default pfall > gpurun_out/r2/ab_ax.log 2>&1
cat gpurun_out/r2/ab_ax.log
