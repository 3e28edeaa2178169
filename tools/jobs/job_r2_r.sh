set -x
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest_r.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest_r.log
tail -4 gpurun_out/r2/pytest_r.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r2/bench_r.json 2> gpurun_out/r2/bench_r.err; echo "bench rc=$?"
grep "\[bench\]" gpurun_out/r2/bench_r.err | tail -3
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2/bench_r_ref.json 2>/dev/null; echo "ref rc=$?"
timeout 120 python tests/fuzz_gpu.py 60 11 > gpurun_out/r2/fuzz_r.log 2>&1; tail -2 gpurun_out/r2/fuzz_r.log
