set -x
mkdir -p gpurun_out/r2
(lscpu; echo; numactl -H 2>/dev/null; echo; nproc; cat /sys/fs/cgroup/cpu.max 2>/dev/null; nvidia-smi topo -m; for d in /sys/bus/pci/devices/*; do if [ -f $d/local_cpulist ] && grep -qi 0x10de $d/vendor 2>/dev/null; then echo $d $(cat $d/class) $(cat $d/local_cpulist) numa $(cat $d/numa_node); fi; done; nvidia-smi --query-gpu=index,pci.bus_id --format=csv) > gpurun_out/r2/host.txt 2>&1
g++ -O3 -march=native -pthread -o /tmp/hostbw tools/micro/hostbw.cpp && /tmp/hostbw 256 > gpurun_out/r2/hostbw.txt 2>&1
python tools/pcie_probe.py > gpurun_out/r2/pcie.txt 2>&1
python tools/ncu_target.py --pattern true --B 65536 > gpurun_out/r2/plain_true.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:eval_kernel -s 1 -c 2 -o gpurun_out/r2/prof_true python tools/ncu_target.py --pattern true --B 65536 > gpurun_out/r2/ncu_true.log 2>&1
python tools/ncu_target.py --pattern block --want f,grad,g --B 65536 > gpurun_out/r2/plain_none.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:eval_kernel -s 1 -c 2 -o gpurun_out/r2/prof_none python tools/ncu_target.py --pattern block --want f,grad,g --B 65536 > gpurun_out/r2/ncu_none.log 2>&1
python tools/ncu_target.py --pattern true --B 4096 --launches 20 >> gpurun_out/r2/plain_true.log 2>&1
python tools/ncu_target.py --pattern block --want f,grad,g --B 4096 --launches 20 >> gpurun_out/r2/plain_none.log 2>&1
python tools/ncu_target.py --pattern block --want g --B 65536 >> gpurun_out/r2/plain_none.log 2>&1
ls -la gpurun_out/r2
