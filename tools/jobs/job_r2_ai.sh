set -x
mkdir -p gpurun_out/r2
# final build: smoke, default bench (both arms), launch list and the three ncu captures, fuzz
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r2/bench_ai.json 2> gpurun_out/r2/bench_ai.err; echo "bench rc=$?"
grep "\[bench\]" gpurun_out/r2/bench_ai.err | tail -4
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2/bench_ai_ref.json 2>/dev/null; echo "ref rc=$?"
python bench.py --steps 20 --warmup 3 --no-extras --no-cpu > gpurun_out/r2/plain_ai.json 2> gpurun_out/r2/plain_ai.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2/launches_ai.csv python bench.py --steps 20 --warmup 3 --no-extras --no-cpu > gpurun_out/r2/ncu_launches_ai.log 2>&1
python bench.py --steps 3 --warmup 3 --no-extras --no-cpu > gpurun_out/r2/plain_ai2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:eval_kernel -s 4 -c 2 -o gpurun_out/r2/prof_block_ai python bench.py --steps 3 --warmup 3 --no-extras --no-cpu > gpurun_out/r2/ncu_block_ai.log 2>&1
python tools/ncu_target.py --pattern true --B 65536 > gpurun_out/r2/plain_true_ai.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:eval_kernel -s 1 -c 1 -o gpurun_out/r2/prof_true_ai python tools/ncu_target.py --pattern true --B 65536 > gpurun_out/r2/ncu_true_ai.log 2>&1
python tools/ncu_target.py --pattern block --want f,grad,g --B 65536 > gpurun_out/r2/plain_none_ai.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:eval_kernel -s 1 -c 1 -o gpurun_out/r2/prof_none_ai python tools/ncu_target.py --pattern block --want f,grad,g --B 65536 > gpurun_out/r2/ncu_none_ai.log 2>&1
timeout 120 python tests/fuzz_gpu.py 60 23 > gpurun_out/r2/fuzz_ai.log 2>&1; tail -2 gpurun_out/r2/fuzz_ai.log
ls -la gpurun_out/r2/*_ai*
python tools/hess_bench.py > gpurun_out/r2/hess_ai.log 2>&1; cat gpurun_out/r2/hess_ai.log
