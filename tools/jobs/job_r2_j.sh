set -x
mkdir -p gpurun_out/r2
nvidia-smi -L | wc -l; nproc
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 50 --warmup 5 > gpurun_out/r2/bench_j_n8.json 2> gpurun_out/r2/bench_j_n8.err; echo "rc=$?"
grep "\[bench\]" gpurun_out/r2/bench_j_n8.err; tail -3 gpurun_out/r2/bench_j_n8.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus 8 --steps 5 --warmup 1 > gpurun_out/r2/bench_j_ref_n8.json 2>/dev/null; echo "rc=$?"
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi_device or current_device" > gpurun_out/r2/pytest_j.log 2>&1; tail -2 gpurun_out/r2/pytest_j.log
g++ -O3 -march=native -pthread -o /tmp/hostbw tools/micro/hostbw.cpp && /tmp/hostbw 128 8 16 32 > gpurun_out/r2/hostbw_j.txt 2>&1; cat gpurun_out/r2/hostbw_j.txt
