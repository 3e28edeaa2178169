set -x
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest_i.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest_i.log
tail -5 gpurun_out/r2/pytest_i.log
timeout 200 python tools/hess_bench.py > gpurun_out/r2/hess_i.log 2>&1; cat gpurun_out/r2/hess_i.log
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/r2/bench_i.json 2> gpurun_out/r2/bench_i.err; echo "bench rc=$?"
grep "\[bench\]" gpurun_out/r2/bench_i.err
