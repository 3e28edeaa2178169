set -x
mkdir -p gpurun_out/r2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r2/bench_ar.json 2> gpurun_out/r2/bench_ar.err; echo "bench rc=$?"
grep "\[bench\]" gpurun_out/r2/bench_ar.err | tail -6
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2/bench_ar_ref.json 2>/dev/null; echo "ref rc=$?"
python bench.py --steps 20 --warmup 3 --no-extras --no-cpu > gpurun_out/r2/plain_ar.json 2> gpurun_out/r2/plain_ar.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2/launches_ar.csv python bench.py --steps 20 --warmup 3 --no-extras --no-cpu > gpurun_out/r2/ncu_launches_ar.log 2>&1
python bench.py --steps 3 --warmup 3 --no-extras --no-cpu > gpurun_out/r2/plain_ar2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:eval_kernel -s 4 -c 2 -o gpurun_out/r2/prof_block_ar python bench.py --steps 3 --warmup 3 --no-extras --no-cpu > gpurun_out/r2/ncu_block_ar.log 2>&1
python tools/ncu_target.py --pattern true --B 65536 > gpurun_out/r2/plain_true_ar.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:eval_kernel -s 1 -c 1 -o gpurun_out/r2/prof_true_ar python tools/ncu_target.py --pattern true --B 65536 > gpurun_out/r2/ncu_true_ar.log 2>&1
python tools/ncu_target.py --pattern block --want f,grad,g --B 65536 > gpurun_out/r2/plain_none_ar.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:eval_kernel -s 1 -c 1 -o gpurun_out/r2/prof_none_ar python tools/ncu_target.py --pattern block --want f,grad,g --B 65536 > gpurun_out/r2/ncu_none_ar.log 2>&1
timeout 100 python tests/fuzz_gpu.py 45 31 > gpurun_out/r2/fuzz_ar.log 2>&1; tail -1 gpurun_out/r2/fuzz_ar.log
