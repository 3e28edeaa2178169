set -x
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest_d.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest_d.log
tail -5 gpurun_out/r2/pytest_d.log
for v in main true10 carve; do
  export QLNLP_LIB=; unset QLNLP_LIB; unset QLNLP_MAX_CARVEOUT
  [ $v = true10 ] && export QLNLP_LIB=$PWD/quadruped_landing_b200/libqlnlp_true10.so
  [ $v = carve ] && export QLNLP_MAX_CARVEOUT=1
  for a in "--pattern true" "--pattern block --want f,grad,g" "--pattern block --want g" "--pattern block"; do
    echo "variant $v" >> gpurun_out/r2/kern_d.log
    timeout 120 python tools/ncu_target.py $a --B 65536 --launches 10 >> gpurun_out/r2/kern_d.log 2>&1
    timeout 120 python tools/ncu_target.py $a --B 4096 --launches 50 >> gpurun_out/r2/kern_d.log 2>&1
  done
done
unset QLNLP_LIB; unset QLNLP_MAX_CARVEOUT
cat gpurun_out/r2/kern_d.log
ALLOC=torch timeout 600 python tools/e2e_probe.py > gpurun_out/r2/e2e_d.log 2>&1
cat gpurun_out/r2/e2e_d.log
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/r2/bench_d.json 2> gpurun_out/r2/bench_d.err; echo "bench rc=$?"
grep "\[bench\]" gpurun_out/r2/bench_d.err
