set -x
mkdir -p gpurun_out/r2
# (compute-sanitizer is closed on this pool; the case still runs every kernel once on small inputs with its own checks)
timeout 300 python tools/sanitize_case.py > gpurun_out/r2/san_plain.log 2>&1; echo "plain rc=$?" | tee -a gpurun_out/r2/san_plain.log
tail -3 gpurun_out/r2/san_plain.log
