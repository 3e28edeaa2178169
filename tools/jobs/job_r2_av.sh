set -x
mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests -m gpu -x -q -k "hessian or Hessian or hess or memory_back or graph" > gpurun_out/r2/pytest_av.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest_av.log
tail -3 gpurun_out/r2/pytest_av.log
python tools/hess_bench.py > gpurun_out/r2/hess_av.log 2>&1; cat gpurun_out/r2/hess_av.log
ncu --set full --clock-control none --import-source on -k regex:hess_kernel -s 8 -c 1 -o gpurun_out/r2/prof_hess_av python tools/hess_bench.py > gpurun_out/r2/ncu_hess_av.log 2>&1
ls -la gpurun_out/r2/prof_hess_av.ncu-rep
