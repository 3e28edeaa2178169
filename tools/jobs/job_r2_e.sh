set -x
mkdir -p gpurun_out/r2
for v in main nopf; do
  unset QLNLP_LIB
  [ $v = nopf ] && export QLNLP_LIB=$PWD/quadruped_landing_b200/libqlnlp_nopf.so
  for a in "--pattern true" "--pattern block --want f,grad,g" "--pattern block --want g" "--pattern block --want g,grad"; do
    echo "variant $v" >> gpurun_out/r2/kern_e.log
    timeout 120 python tools/ncu_target.py $a --B 65536 --launches 10 >> gpurun_out/r2/kern_e.log 2>&1
    timeout 120 python tools/ncu_target.py $a --B 4096 --launches 50 >> gpurun_out/r2/kern_e.log 2>&1
  done
done
unset QLNLP_LIB
cat gpurun_out/r2/kern_e.log | cut -c1-100
PER_CALL=1 CHUNKS=64,96,128,256,384,512,1024 ALLOC=torch timeout 600 python tools/e2e_probe.py > gpurun_out/r2/e2e_e_torch.log 2>&1
PER_CALL=1 CHUNKS=64,128,256,512 ALLOC=huge timeout 600 python tools/e2e_probe.py > gpurun_out/r2/e2e_e_huge.log 2>&1
grep -v "per call\|calls ms" gpurun_out/r2/e2e_e_torch.log gpurun_out/r2/e2e_e_huge.log
