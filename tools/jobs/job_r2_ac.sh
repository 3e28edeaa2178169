set -x
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2/pytest_ac.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest_ac.log
tail -3 gpurun_out/r2/pytest_ac.log
timeout 1200 python tools/ab_bench.py run head default default@QLNLP_MAX_CARVEOUT=1 head default default@QLNLP_MAX_CARVEOUT=1 > gpurun_out/r2/ab_ac.log 2>&1
cat gpurun_out/r2/ab_ac.log
python tools/ncu_target.py --pattern block --want f,grad,g --B 65536 > /dev/null 2>&1 &&
ncu --metrics launch__shared_mem_config_size,gpu__time_duration.sum -k regex:eval_kernel -c 2 --csv --log-file gpurun_out/r2/cfg_none_ac.csv python tools/ncu_target.py --pattern block --want f,grad,g --B 65536 > /dev/null 2>&1
cat gpurun_out/r2/cfg_none_ac.csv | tail -5
