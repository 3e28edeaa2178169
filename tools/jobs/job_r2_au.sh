set -x
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest_au.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest_au.log
tail -3 gpurun_out/r2/pytest_au.log
for i in 1 2 3; do timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stream_order" 2>&1 | tail -1; done
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
