set -x
mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests -m gpu -x -q -k "hessian or Hessian or hess or graph" > gpurun_out/r2/pytest_aw.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest_aw.log
tail -2 gpurun_out/r2/pytest_aw.log
for i in 1 2; do python tools/hess_bench.py 2>&1 | tee -a gpurun_out/r2/hess_aw.log; done
