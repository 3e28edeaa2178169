set -x
mkdir -p gpurun_out/r2
for d in 1 0 1 0; do
  echo "=== direct $d" >> gpurun_out/r2/e2e_t.log
  QLNLP_HOST_DIRECT=$d CHUNKS=512 ALLOC=huge timeout 300 python tools/e2e_probe.py 2>&1 | grep "registered=1 chunk\|registered=0 chunk\|per call\|want=" | head -12 >> gpurun_out/r2/e2e_t.log
done
cat gpurun_out/r2/e2e_t.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "host or multi_device or empty" > gpurun_out/r2/pytest_t.log 2>&1; tail -3 gpurun_out/r2/pytest_t.log
