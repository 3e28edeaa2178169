set -x
mkdir -p gpurun_out/r2
QLNLP_HOST_SUBCHUNK=48 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "host_batch or host_pointer or multi_device or empty" > gpurun_out/r2/pytest_k.log 2>&1; tail -3 gpurun_out/r2/pytest_k.log
for cfg in "0 3" "32 2" "32 3" "64 2" "64 3" "64 4" "128 3" "128 4" "0 3"; do
  set -- $cfg
  echo "=== subchunk $1 ring $2" >> gpurun_out/r2/e2e_k.log
  QLNLP_HOST_SUBCHUNK=$1 QLNLP_HOST_RING=$2 CHUNKS=512 ALLOC=torch timeout 300 python tools/e2e_probe.py 2>&1 | grep "block registered=1 chunk\|block registered=0 chunk\|per call" | head -4 >> gpurun_out/r2/e2e_k.log
done
cat gpurun_out/r2/e2e_k.log
for w in 5 6 7 8; do echo "warps/SM $w" >> gpurun_out/r2/kern_k.log; QLNLP_BLOCKS_PER_SM=$w python tools/ncu_target.py --pattern block --B 4096 --launches 100 >> gpurun_out/r2/kern_k.log 2>&1;  QLNLP_BLOCKS_PER_SM=$w python tools/ncu_target.py --pattern block --want g,jac --B 4096 --launches 100 >> gpurun_out/r2/kern_k.log 2>&1; done
python tools/ncu_target.py --pattern block --B 4096 --launches 100 >> gpurun_out/r2/kern_k.log 2>&1
cut -c1-110 gpurun_out/r2/kern_k.log
