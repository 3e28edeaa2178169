set -x
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest_a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest_a.log
tail -5 gpurun_out/r2/pytest_a.log
THREADS_SWEEP=1 timeout 600 python tools/e2e_probe.py > gpurun_out/r2/e2e_a.log 2>&1
timeout 300 python tools/single_eval_latency.py > gpurun_out/r2/single_a.log 2>&1
for a in "--pattern block" "--pattern true" "--pattern block --want f,grad,g" "--pattern block --want g" "--pattern block --want g,jac"; do
  timeout 120 python tools/ncu_target.py $a --B 65536 --launches 10 >> gpurun_out/r2/kern_a.log 2>&1
  timeout 120 python tools/ncu_target.py $a --B 4096 --launches 50 >> gpurun_out/r2/kern_a.log 2>&1
done
cat gpurun_out/r2/kern_a.log
