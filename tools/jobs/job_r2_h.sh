set -x
mkdir -p gpurun_out/r2
nvidia-smi -L; nproc
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi_device or current_device or host_batch" > gpurun_out/r2/pytest_h.log 2>&1; tail -3 gpurun_out/r2/pytest_h.log
python tests/two_device_check.py > gpurun_out/r2/two_dev_h.log 2>&1; tail -2 gpurun_out/r2/two_dev_h.log
DEVICES=0,1 B=8192 CHUNKS=512 timeout 300 python tools/e2e_probe.py > gpurun_out/r2/e2e_h_multi.log 2>&1; grep -v "per call" gpurun_out/r2/e2e_h_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/r2/bench_h_n2.json 2> gpurun_out/r2/bench_h_n2.err; echo "rc=$?"
grep "\[bench\]" gpurun_out/r2/bench_h_n2.err; tail -3 gpurun_out/r2/bench_h_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 5 --warmup 1 > gpurun_out/r2/bench_h_ref_n2.json 2>/dev/null; echo "rc=$?"
