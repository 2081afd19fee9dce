#!/bin/bash
# A/B of scoring-kernel builds on one GPU box: for every variant (extra nvcc flags) rebuild the
# library, run the parity tests that pin the fast kernel, then time it (stage times of bench.py).
#   gpurun --timeout 1500 -- 'bash tools/k2_variants.sh "" "-DK2_SWEEP=4" "-DK2_MIN_CTAS=5"'
set -u
mkdir -p gpurun_out
: > gpurun_out/k2_variants.txt
for v in "$@"; do
    tag=$(echo "default $v" | tr -c 'A-Za-z0-9_=\n' '_')
    TSK_NVCC_EXTRA="$v" python -m tailseeker_b200.build --force > /dev/null || { echo "$v: build failed" >> gpurun_out/k2_variants.txt; continue; }
    python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -x -q -m gpu > gpurun_out/k2_tests_$tag.log 2>&1
    st=$?
    python bench.py --tiles 4 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/k2_bench_$tag.json 2> gpurun_out/k2_bench_$tag.err
    python - "$v" "$st" gpurun_out/k2_bench_$tag.json >> gpurun_out/k2_variants.txt <<'PY'
import json, sys
v, st, path = sys.argv[1:4]
try:
    b = json.loads(open(path).read().strip().splitlines()[-1])
    s = b["roofline"]["stages_ms_per_tile"]
    print("%-40s tests %s  K2 %.3f ms  decode %.3f  ruler %.3f  fixups %.3f  -> %.1f M clusters/s" % (
        v or "(default)", "ok" if st == "0" else "FAILED", s["normalise_score_pack"], s["decode_demux_seed"],
        s["polya_ruler"], s["fair_sampling_fixups"], b["value"] / 1e6))
except Exception as e:
    print("%-40s tests %s  bench failed: %s" % (v or "(default)", st, e))
PY
done
python -m tailseeker_b200.build --force > /dev/null
cat gpurun_out/k2_variants.txt
