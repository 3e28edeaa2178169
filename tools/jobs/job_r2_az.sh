set -x
mkdir -p gpurun_out/r2
run() { tag=$1; shift; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --no-cpu --steps 100 --warmup 10 > gpurun_out/r2/bench_az_$tag.json 2>/dev/null
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2/bench_az_$tag.json") if l.startswith("{")][-1])
print("$tag", "value/GPU", round(d["value"]/d["n_gpus"]/1e6,3))
PY
}
run full A=1
run skip_c3 QL_BENCH_SKIP=c3
run skip_c4c5 QL_BENCH_SKIP=c4,c5
run skip_variants QL_BENCH_SKIP=variants
