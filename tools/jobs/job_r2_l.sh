set -x
mkdir -p gpurun_out/r2
for r in 1 0 1 0; do
  echo "=== ramp $r" >> gpurun_out/r2/e2e_l.log
  QLNLP_HOST_RAMP=$r CHUNKS=512 ALLOC=huge timeout 300 python tools/e2e_probe.py 2>&1 | grep "registered=1 chunk\|registered=0 chunk\|per call" | head -8 >> gpurun_out/r2/e2e_l.log
done
cat gpurun_out/r2/e2e_l.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest_l.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest_l.log
tail -4 gpurun_out/r2/pytest_l.log
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/r2/bench_l.json 2> gpurun_out/r2/bench_l.err; echo "bench rc=$?"
grep "\[bench\]" gpurun_out/r2/bench_l.err
