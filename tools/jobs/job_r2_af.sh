set -x
mkdir -p gpurun_out/r2
timeout 1200 python tools/ab_bench.py run default nw14 nw12 tw7 default nw14 nw12 tw7 > gpurun_out/r2/ab_af.log 2>&1
cat gpurun_out/r2/ab_af.log
