set -x
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest_am.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest_am.log
tail -3 gpurun_out/r2/pytest_am.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r2/bench_am.json 2> gpurun_out/r2/bench_am.err; echo "bench rc=$?"
grep "\[bench\]" gpurun_out/r2/bench_am.err | tail -4
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2/bench_am_ref.json 2>/dev/null; echo "ref rc=$?"
python tools/single_eval_latency.py > gpurun_out/r2/single_am.log 2>&1; tail -8 gpurun_out/r2/single_am.log
