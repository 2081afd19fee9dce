#!/bin/bash
# First GPU call of the next round: measures the default build (single-correction division, never timed
# so far), then the (k, i)-table logf of -DTSK_LOGF_KI=1 (never run so far) -- parity first, time second.
#   gpurun --timeout 1500 -- 'bash tools/try_logf_ki.sh'
# Results land in gpurun_out/: bench_default.json, ki_tests.log, bench_ki.json.
set -u
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
TSK_NVCC_EXTRA=-DTSK_LOGF_KI=1 python -m tailseeker_b200.build --force || exit 1
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_golden.py -x -q -m gpu \
    > gpurun_out/ki_tests.log 2>&1
status=$?
tail -3 gpurun_out/ki_tests.log
if [ $status -eq 0 ]; then
    python bench.py --steps 3 --warmup 3 > gpurun_out/bench_ki.json 2> gpurun_out/bench_ki.err
    python - <<'PY'
import json
a = json.loads(open("gpurun_out/bench_default.json").read().strip().splitlines()[-1])
b = json.loads(open("gpurun_out/bench_ki.json").read().strip().splitlines()[-1])
print("default  %.1f M clusters/s, K2 roofline %.3f" % (a["value"] / 1e6, a["roofline"]["frac"]))
print("LOGF_KI  %.1f M clusters/s, K2 roofline %.3f" % (b["value"] / 1e6, b["roofline"]["frac"]))
PY
fi
python -m tailseeker_b200.build --force          # back to the default build
exit $status
