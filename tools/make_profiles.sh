#!/bin/bash
# Regenerate profiles/r02_eval_kernel_{block,true,none}.md (+ traffic.json, launch list) from the .ncu-rep files a
# gpurun job brought back:   tools/make_profiles.sh <tag>     (reads gpurun_out/r2/prof_{block,true,none}_<tag>.ncu-rep)
set -e
cd "$(dirname "$0")/.."
tag=$1
hdr='## hottest CUDA source lines (first launch; tools/ncu_lines.py: share of stall samples, share of executed warp instructions, excess shared-memory wavefronts, top stall reasons)'
python tools/ncu_summary.py gpurun_out/r2/prof_block_$tag.ncu-rep profiles/r02_eval_kernel_block --traffic > /dev/null
python tools/ncu_summary.py gpurun_out/r2/prof_true_$tag.ncu-rep profiles/r02_eval_kernel_true \
    --what "tools/ncu_target.py --pattern true --B 65536 (SPARSE_TRUE pattern, f+grad+g+J, one launch of 65,536 evaluations)" > /dev/null
python tools/ncu_summary.py gpurun_out/r2/prof_none_$tag.ncu-rep profiles/r02_eval_kernel_none \
    --what "tools/ncu_target.py --pattern block --want f,grad,g --B 65536 (no Jacobian: f+grad+g, one launch of 65,536 evaluations)" > /dev/null
for k in block true none; do
    { echo; echo "$hdr"; echo '```'; python tools/ncu_lines.py gpurun_out/r2/prof_${k}_$tag.ncu-rep 26; echo '```'; } >> profiles/r02_eval_kernel_$k.md
done
cp gpurun_out/r2/launches_$tag.csv profiles/r02_launches.csv
