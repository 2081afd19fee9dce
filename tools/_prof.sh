CMD="python bench.py --tiles 2 --steps 1 --no-e2e --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_|^void k_" -s 60 -c 60 --csv --log-file gpurun_out/r1_launches.csv $CMD > gpurun_out/ncu_l.log 2>&1
