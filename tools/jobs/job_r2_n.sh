set -x
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest_n.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest_n.log
tail -6 gpurun_out/r2/pytest_n.log
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/r2/bench_n.json 2> gpurun_out/r2/bench_n.err; echo "bench rc=$?"
grep "\[bench\]" gpurun_out/r2/bench_n.err; tail -2 gpurun_out/r2/bench_n.err
