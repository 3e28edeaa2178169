set -x
mkdir -p gpurun_out/r2
timeout 1200 python tools/ab_bench.py run head default default@QLNLP_MAX_CARVEOUT=1 head default default@QLNLP_MAX_CARVEOUT=1 default > gpurun_out/r2/ab_ae.log 2>&1
cat gpurun_out/r2/ab_ae.log
for k in "--pattern block --want f,grad,g" "--pattern true" "--pattern block"; do
ncu --metrics launch__shared_mem_config_size -k regex:eval_kernel -c 1 --csv --log-file gpurun_out/r2/cfg_ae.csv python tools/ncu_target.py $k --B 65536 > /dev/null 2>&1
echo "$k: $(grep shared_mem_config gpurun_out/r2/cfg_ae.csv | awk -F, '{print $5, $NF}')" | tee -a gpurun_out/r2/cfg_ae.log
done
