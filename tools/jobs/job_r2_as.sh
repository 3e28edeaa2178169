set -x
mkdir -p gpurun_out/r2
for i in 1 2 3; do
for vf in 0 1; do
echo "== VALS_FIRST=$vf run $i" >> gpurun_out/r2/e2e_as.log
QLNLP_HOST_VALS_FIRST=$vf ALLOC=huge CHUNKS=512 python tools/e2e_probe.py 2>&1 | grep -v "^output" >> gpurun_out/r2/e2e_as.log
done; done
cat gpurun_out/r2/e2e_as.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "host" > gpurun_out/r2/pytest_as.log 2>&1; tail -2 gpurun_out/r2/pytest_as.log
