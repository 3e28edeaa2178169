set -x
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest_m.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest_m.log
tail -15 gpurun_out/r2/pytest_m.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2/smoke_m.log 2>&1; tail -2 gpurun_out/r2/smoke_m.log
