set -x
mkdir -p gpurun_out/r2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 100 --warmup 10 > gpurun_out/r2/bench_ak_n8.json 2> gpurun_out/r2/bench_ak_n8.err; echo "rc=$?"
grep "\[bench\]" gpurun_out/r2/bench_ak_n8.err | tail -4; tail -2 gpurun_out/r2/bench_ak_n8.err | cut -c1-200
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus 8 --steps 10 --warmup 2 > gpurun_out/r2/bench_ak_ref_n8.json 2>/dev/null; echo "rc=$?"
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi_device" > gpurun_out/r2/pytest_ak.log 2>&1; tail -2 gpurun_out/r2/pytest_ak.log
