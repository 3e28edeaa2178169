set -x
mkdir -p gpurun_out/r2
for g in 0 888 820 740 683 586 512; do
  echo "grid $g" >> gpurun_out/r2/kern_p.log
  if [ $g = 0 ]; then unset QLNLP_GRID; else export QLNLP_GRID=$g; fi
  python tools/ncu_target.py --pattern block --B 4096 --launches 100 >> gpurun_out/r2/kern_p.log 2>&1
  python tools/ncu_target.py --pattern block --want g,jac --B 4096 --launches 100 >> gpurun_out/r2/kern_p.log 2>&1
done
unset QLNLP_GRID
for a in "--pattern true" "--pattern block --want f,grad,g" "--pattern block --want g"; do
    timeout 120 python tools/ncu_target.py $a --B 65536 --launches 10 >> gpurun_out/r2/kern_p.log 2>&1
done
cut -c1-100 gpurun_out/r2/kern_p.log
