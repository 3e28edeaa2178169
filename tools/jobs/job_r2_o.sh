set -x
mkdir -p gpurun_out/r2
for v in main prev main prev; do
  unset QLNLP_LIB
  [ $v = prev ] && export QLNLP_LIB=$PWD/quadruped_landing_b200/libqlnlp_prev.so
  for a in "--pattern true" "--pattern block --want f,grad,g" "--pattern block --want g" "--pattern block"; do
    echo "variant $v" >> gpurun_out/r2/kern_o.log
    timeout 120 python tools/ncu_target.py $a --B 65536 --launches 10 >> gpurun_out/r2/kern_o.log 2>&1
    timeout 120 python tools/ncu_target.py $a --B 4096 --launches 50 >> gpurun_out/r2/kern_o.log 2>&1
  done
done
unset QLNLP_LIB
cut -c1-100 gpurun_out/r2/kern_o.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest_o.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest_o.log
tail -4 gpurun_out/r2/pytest_o.log
