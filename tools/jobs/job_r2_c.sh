set -x
mkdir -p gpurun_out/r2
python tools/hostrows_bench.py 2>&1 | grep -v numpy > gpurun_out/r2/hostrows_c.log
ALLOC=torch timeout 600 python tools/e2e_probe.py > gpurun_out/r2/e2e_c_torch.log 2>&1
QLNLP_ONE_ZEROCOPY=1 timeout 300 python tools/single_eval_latency.py > gpurun_out/r2/single_c_zc1.log 2>&1
QLNLP_ONE_ZEROCOPY=0 timeout 300 python tools/single_eval_latency.py > gpurun_out/r2/single_c_zc0.log 2>&1
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/r2/bench_c.json 2> gpurun_out/r2/bench_c.err; echo "bench rc=$?"
tail -5 gpurun_out/r2/bench_c.err
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "host or eval_all or known" > gpurun_out/r2/pytest_c.log 2>&1; tail -3 gpurun_out/r2/pytest_c.log
