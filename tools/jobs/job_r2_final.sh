set -x
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest_final.log
tail -3 gpurun_out/r2/pytest_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r2/bench_final.json 2> gpurun_out/r2/bench_final.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2/bench_final_ref.json 2>/dev/null; echo "ref rc=$?"
