set -x
mkdir -p gpurun_out/r2
python tools/ncu_target.py --pattern block --want f,grad,g --B 65536 > /dev/null 2>&1
for pct in 50 58 65 72 79 86 90 95 100; do
  QLNLP_CARVEOUT_PCT=$pct ncu --metrics launch__shared_mem_config_size,gpu__time_duration.sum -k regex:eval_kernel -c 1 --csv --log-file gpurun_out/r2/cfg_$pct.csv python tools/ncu_target.py --pattern block --want f,grad,g --B 65536 > /dev/null 2>&1
  echo "pct=$pct: $(grep shared_mem_config gpurun_out/r2/cfg_$pct.csv | awk -F, '{print $NF}')  $(grep time_duration gpurun_out/r2/cfg_$pct.csv | awk -F, '{print $NF}')" | tee -a gpurun_out/r2/cfg_scan.log
done
for pct in 72 86; do
  QLNLP_CARVEOUT_PCT=$pct ncu --metrics launch__shared_mem_config_size,gpu__time_duration.sum -k regex:eval_kernel -c 1 --csv --log-file gpurun_out/r2/cfgt_$pct.csv python tools/ncu_target.py --pattern true --B 65536 > /dev/null 2>&1
  echo "TRUE pct=$pct: $(grep shared_mem_config gpurun_out/r2/cfgt_$pct.csv | awk -F, '{print $NF}')" | tee -a gpurun_out/r2/cfg_scan.log
done
