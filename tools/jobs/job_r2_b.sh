set -x
mkdir -p gpurun_out/r2
python tools/hostrows_bench.py > gpurun_out/r2/hostrows_b.log 2>&1
ALLOC=torch timeout 600 python tools/e2e_probe.py > gpurun_out/r2/e2e_b_torch.log 2>&1
ALLOC=huge timeout 600 python tools/e2e_probe.py > gpurun_out/r2/e2e_b_huge.log 2>&1
timeout 300 python tools/single_eval_latency.py > gpurun_out/r2/single_b.log 2>&1
python tools/ncu_target.py --pattern block --want f,grad,g --B 65536 > gpurun_out/r2/plain_none_b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:eval_kernel -s 1 -c 1 -o gpurun_out/r2/prof_none_b python tools/ncu_target.py --pattern block --want f,grad,g --B 65536 > gpurun_out/r2/ncu_none_b.log 2>&1
python tools/ncu_target.py --pattern true --B 65536 > gpurun_out/r2/plain_true_b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:eval_kernel -s 1 -c 1 -o gpurun_out/r2/prof_true_b python tools/ncu_target.py --pattern true --B 65536 > gpurun_out/r2/ncu_true_b.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest_b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest_b.log
tail -5 gpurun_out/r2/pytest_b.log
