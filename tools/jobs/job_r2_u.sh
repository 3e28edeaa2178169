set -x
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest_u.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest_u.log
tail -4 gpurun_out/r2/pytest_u.log
python bench.py --steps 20 --warmup 3 --no-extras --no-cpu > gpurun_out/r2/plain_u.json 2> gpurun_out/r2/plain_u.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2/launches_u.csv python bench.py --steps 20 --warmup 3 --no-extras --no-cpu > gpurun_out/r2/ncu_launches_u.log 2>&1
python bench.py --steps 3 --warmup 3 --no-extras --no-cpu > gpurun_out/r2/plain_u2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:eval_kernel -s 4 -c 2 -o gpurun_out/r2/prof_block_u python bench.py --steps 3 --warmup 3 --no-extras --no-cpu > gpurun_out/r2/ncu_block_u.log 2>&1
python tools/ncu_target.py --pattern true --B 65536 > gpurun_out/r2/plain_true_u.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:eval_kernel -s 1 -c 1 -o gpurun_out/r2/prof_true_u python tools/ncu_target.py --pattern true --B 65536 > gpurun_out/r2/ncu_true_u.log 2>&1
python tools/ncu_target.py --pattern block --want f,grad,g --B 65536 > gpurun_out/r2/plain_none_u.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:eval_kernel -s 1 -c 1 -o gpurun_out/r2/prof_none_u python tools/ncu_target.py --pattern block --want f,grad,g --B 65536 > gpurun_out/r2/ncu_none_u.log 2>&1
python tools/hess_bench.py > gpurun_out/r2/hess_u.log 2>&1; cat gpurun_out/r2/hess_u.log
ls gpurun_out/r2/*_u*
