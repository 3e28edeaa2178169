set -x
mkdir -p gpurun_out/r2
python bench.py --no-extras --no-cpu > gpurun_out/r2/bench_ay_n1.json 2>/dev/null
CUDA_VISIBLE_DEVICES=1 python bench.py --no-extras --no-cpu > gpurun_out/r2/bench_ay_n1b.json 2>/dev/null
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --no-extras --no-cpu > gpurun_out/r2/bench_ay_n2.json 2>/dev/null
QLNLP_PDL=0 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 2 --no-extras --no-cpu > gpurun_out/r2/bench_ay_n2_nopdl.json 2>/dev/null
python - <<'PY'
import json
for f in ("n1","n1b","n2","n2_nopdl"):
    d=json.loads([l for l in open(f"gpurun_out/r2/bench_ay_{f}.json") if l.startswith("{")][-1])
    print(f, "value/GPU", round(d["value"]/d["n_gpus"]/1e6,3), "ms_per_step", round(d["ms_per_step"],4))
PY
