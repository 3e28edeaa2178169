set -x
mkdir -p gpurun_out/r2
nvidia-smi --query-gpu=name,clocks.sm,clocks.mem,temperature.gpu,power.draw --format=csv
timeout 1200 python tools/ab_bench.py run default base default base mra default base mra > gpurun_out/r2/ab_y.log 2>&1
cat gpurun_out/r2/ab_y.log
