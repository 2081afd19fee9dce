#!/bin/bash
# Round-2 profile evidence (one GPU): launch list of a short bench run, ncu --set full of the five
# kernels that matter (steady state; both instances of the scoring kernel), results in gpurun_out/;
# tools/profile_digest.py then writes the tracked summaries under profiles/.
#   gpurun --timeout 1800 -- 'bash tools/profile_r2.sh'
set -u
mkdir -p gpurun_out
B="python bench.py --tiles 2 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-cli --no-extra"
$B > gpurun_out/prof_bench.json 2> gpurun_out/prof_bench.err || { echo "bench failed"; tail -5 gpurun_out/prof_bench.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file gpurun_out/r2_launches.csv $B > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -f -o gpurun_out/r2_top \
    -k regex:'k_decode_demux_seed|k_polya_ruler|k_normalise_score_pack_compact|k_enc_deflate|k_enc_format' -s 20 -c 8 \
    $B > gpurun_out/ncu_top.log 2>&1
ls -la gpurun_out/r2_top.ncu-rep gpurun_out/r2_launches.csv
