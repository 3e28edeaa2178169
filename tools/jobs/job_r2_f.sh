set -x
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest_f.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest_f.log
tail -5 gpurun_out/r2/pytest_f.log
timeout 300 python tools/ragged_bench.py > gpurun_out/r2/ragged_f.log 2>&1; cat gpurun_out/r2/ragged_f.log
for a in "--pattern true" "--pattern block --want f,grad,g" "--pattern block --want g" "--pattern block"; do
    timeout 120 python tools/ncu_target.py $a --B 65536 --launches 10 >> gpurun_out/r2/kern_f.log 2>&1
    timeout 120 python tools/ncu_target.py $a --B 4096 --launches 50 >> gpurun_out/r2/kern_f.log 2>&1
done
cut -c1-100 gpurun_out/r2/kern_f.log
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/r2/bench_f.json 2> gpurun_out/r2/bench_f.err; echo "bench rc=$?"
grep "\[bench\]" gpurun_out/r2/bench_f.err
