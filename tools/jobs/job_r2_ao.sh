set -x
mkdir -p gpurun_out/r2
for pdl in 1 0 1 0; do
QLNLP_PDL=$pdl python bench.py --no-cpu > gpurun_out/r2/bench_ao_$pdl.json 2>/dev/null
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2/bench_ao_$pdl.json') if l.startswith('{')][-1])
print('PDL=$pdl full bench: value', round(d['value']/1e6,3), 'frac', round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value']/1e3,1), 'c5', round(d['c5']['value']/1e6,2), 'c3', round(d['c3']['value']/1e6,2), 'clocks', d['clocks'])
" | tee -a gpurun_out/r2/bench_ao.log
done
