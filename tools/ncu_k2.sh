#!/bin/bash
# ncu --set full capture of the scoring kernel for one or more builds (extra nvcc flags per argument);
# reports land in gpurun_out/k2_<tag>.ncu-rep.  One GPU, one kernel launch each.
#   gpurun --timeout 1200 -- 'bash tools/ncu_k2.sh "" "-DTSK_K2_LOGTAB_SMEM"'
set -u
mkdir -p gpurun_out
for v in "$@"; do
    tag=$(echo "r2$v" | tr -c 'A-Za-z0-9_=\n' '_')
    TSK_NVCC_EXTRA="$v" python -m tailseeker_b200.build --force > /dev/null || continue
    python bench.py --tiles 2 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2>&1 || { echo "bench failed for $v"; continue; }
    ncu --set full --clock-control none --import-source on -k regex:k_normalise_score_pack_compact -s 2 -c 1 -f \
        -o gpurun_out/k2_$tag python bench.py --tiles 2 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_$tag.log 2>&1
    ls -la gpurun_out/k2_$tag.ncu-rep
done
python -m tailseeker_b200.build --force > /dev/null
