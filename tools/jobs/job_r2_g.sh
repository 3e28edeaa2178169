set -x
mkdir -p gpurun_out/r2
python bench.py --steps 20 --warmup 3 --no-extras --no-cpu > gpurun_out/r2/bench_g_noextras.json 2> gpurun_out/r2/bench_g_noextras.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2/launches_g.csv python bench.py --steps 20 --warmup 3 --no-extras --no-cpu > gpurun_out/r2/ncu_launches_g.log 2>&1
grep "\[bench\]" gpurun_out/r2/bench_g_noextras.err
OMP_WAIT_POLICY=passive python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/r2/bench_g_passive.json 2> gpurun_out/r2/bench_g_passive.err
grep "\[bench\]" gpurun_out/r2/bench_g_passive.err
python bench.py --steps 3 --warmup 3 --no-extras --no-cpu > gpurun_out/r2/plain_g.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:eval_kernel -s 4 -c 2 -o gpurun_out/r2/prof_block_g python bench.py --steps 3 --warmup 3 --no-extras --no-cpu > gpurun_out/r2/ncu_block_g.log 2>&1
python tools/ncu_target.py --pattern true --B 65536 > gpurun_out/r2/plain_true_g.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:eval_kernel -s 1 -c 1 -o gpurun_out/r2/prof_true_g python tools/ncu_target.py --pattern true --B 65536 > gpurun_out/r2/ncu_true_g.log 2>&1
python tools/ncu_target.py --pattern block --want f,grad,g --B 65536 > gpurun_out/r2/plain_none_g.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:eval_kernel -s 1 -c 1 -o gpurun_out/r2/prof_none_g python tools/ncu_target.py --pattern block --want f,grad,g --B 65536 > gpurun_out/r2/ncu_none_g.log 2>&1
ls -la gpurun_out/r2 | tail -12
