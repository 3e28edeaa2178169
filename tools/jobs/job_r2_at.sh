set -x
mkdir -p gpurun_out/r2
N=${NG:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/r2/bench_at_n$N.json 2> gpurun_out/r2/bench_at_n$N.err; echo "rc=$?"
grep "\[bench\]" gpurun_out/r2/bench_at_n$N.err | tail -2
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus $N --steps 10 --warmup 2 > gpurun_out/r2/bench_at_ref_n$N.json 2>/dev/null; echo "rc=$?"
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi_device or current_device" > gpurun_out/r2/pytest_at_n$N.log 2>&1; tail -1 gpurun_out/r2/pytest_at_n$N.log
