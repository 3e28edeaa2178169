set -x
mkdir -p gpurun_out/r2
# final build: smoke, default bench (both arms), launch list and the three ncu captures, fuzz
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r2/bench_aa.json 2> gpurun_out/r2/bench_aa.err; echo "bench rc=$?"
grep "\[bench\]" gpurun_out/r2/bench_aa.err | tail -4
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2/bench_aa_ref.json 2>/dev/null; echo "ref rc=$?"
python bench.py --steps 20 --warmup 3 --no-extras --no-cpu > gpurun_out/r2/plain_aa.json 2> gpurun_out/r2/plain_aa.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2/launches_aa.csv python bench.py --steps 20 --warmup 3 --no-extras --no-cpu > gpurun_out/r2/ncu_launches_aa.log 2>&1
python bench.py --steps 3 --warmup 3 --no-extras --no-cpu > gpurun_out/r2/plain_aa2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:eval_kernel -s 4 -c 2 -o gpurun_out/r2/prof_block_aa python bench.py --steps 3 --warmup 3 --no-extras --no-cpu > gpurun_out/r2/ncu_block_aa.log 2>&1
python tools/ncu_target.py --pattern true --B 65536 > gpurun_out/r2/plain_true_aa.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:eval_kernel -s 1 -c 1 -o gpurun_out/r2/prof_true_aa python tools/ncu_target.py --pattern true --B 65536 > gpurun_out/r2/ncu_true_aa.log 2>&1
python tools/ncu_target.py --pattern block --want f,grad,g --B 65536 > gpurun_out/r2/plain_none_aa.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:eval_kernel -s 1 -c 1 -o gpurun_out/r2/prof_none_aa python tools/ncu_target.py --pattern block --want f,grad,g --B 65536 > gpurun_out/r2/ncu_none_aa.log 2>&1
timeout 120 python tests/fuzz_gpu.py 60 23 > gpurun_out/r2/fuzz_aa.log 2>&1; tail -2 gpurun_out/r2/fuzz_aa.log
ls -la gpurun_out/r2/*_aa*
