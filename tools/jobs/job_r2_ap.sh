set -x
mkdir -p gpurun_out/r2
run() { # label, env..., -- args
  label=$1; shift
  env "$@" python bench.py --no-cpu $EXTRA > gpurun_out/r2/bench_ap.json 2>/dev/null
  python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2/bench_ap.json') if l.startswith('{')][-1])
print('$label: value', round(d['value']/1e6,3), 'frac', round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value']/1e3,1), 'c5', round(d['c5']['value']/1e6,2) if 'c5' in d else '-')
" | tee -a gpurun_out/r2/bench_ap.log
}
EXTRA=--no-extras run "A no-extras PDL=1 MAXCARVE PAD=0 (repro?)" QLNLP_PDL=1 QLNLP_MAX_CARVEOUT=1 QLNLP_PAD_SMEM=0
EXTRA=--no-extras run "B no-extras PDL=1 MAXCARVE PAD=1" QLNLP_PDL=1 QLNLP_MAX_CARVEOUT=1 QLNLP_PAD_SMEM=1
EXTRA= run "C full PDL=1 PAD=1 (new default)" QLNLP_PDL=1
EXTRA= run "D full PDL=1 PAD=0" QLNLP_PDL=1 QLNLP_PAD_SMEM=0
EXTRA= run "E full PDL=0 PAD=1" QLNLP_PDL=0
EXTRA= run "C2 full default" QLNLP_PDL=1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2/pytest_ap.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest_ap.log
tail -3 gpurun_out/r2/pytest_ap.log
