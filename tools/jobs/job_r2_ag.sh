set -x
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest_ag.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest_ag.log
tail -3 gpurun_out/r2/pytest_ag.log
python tools/ncu_target.py --pattern block --want f,grad,g --B 65536 > gpurun_out/r2/plain_none_ag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:eval_kernel -s 1 -c 1 -o gpurun_out/r2/prof_none_ag python tools/ncu_target.py --pattern block --want f,grad,g --B 65536 > gpurun_out/r2/ncu_none_ag.log 2>&1
python tools/ncu_target.py --pattern true --B 65536 > gpurun_out/r2/plain_true_ag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:eval_kernel -s 1 -c 1 -o gpurun_out/r2/prof_true_ag python tools/ncu_target.py --pattern true --B 65536 > gpurun_out/r2/ncu_true_ag.log 2>&1
ls gpurun_out/r2/*_ag*
