set -x
mkdir -p gpurun_out/r2
for i in 1 2; do
for pdl in 0 1; do
QLNLP_PDL=$pdl python bench.py --no-extras --no-cpu > gpurun_out/r2/bench_an_$pdl$i.json 2>/dev/null
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2/bench_an_$pdl$i.json') if l.startswith('{')][-1])
print('PDL=$pdl run $i: value', round(d['value']/1e6,3), 'frac', round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value']/1e3,1), 'steps', d['steps'])
" | tee -a gpurun_out/r2/bench_an.log
done; done
for pdl in 0 1; do
QLNLP_PDL=$pdl python bench.py --no-extras --no-cpu --steps 50 --warmup 10 > gpurun_out/r2/bench_an_s$pdl.json 2>/dev/null
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2/bench_an_s$pdl.json') if l.startswith('{')][-1])
print('PDL=$pdl steps 50: value', round(d['value']/1e6,3), 'frac', round(d['roofline']['frac'],4))
" | tee -a gpurun_out/r2/bench_an.log
done
