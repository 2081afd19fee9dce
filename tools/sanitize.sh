#!/bin/bash
# compute-sanitizer over the hot path on a small tile (smoke(): 3000 clusters through every kernel of the
# importer, the ruler, the dedup kernels) and over the device-side encoder; logs land in gpurun_out/.
#   gpurun --timeout 1500 -- 'bash tools/sanitize.sh'
set -u
mkdir -p gpurun_out
cat > /tmp/tsk_san.py <<'PY'
import sys
sys.path.insert(0, ".")
import numpy as np
import __graft_entry__ as g
g.smoke()
from tailseeker_b200.config import stock_config
from tailseeker_b200.importer import TileProcessor, invert_colormatrix
from tailseeker_b200.synth import make_tile
cfg = stock_config("miseq", phix_match_name="PhiX")
t = make_tile(cfg, 2000, seed=3, dark_cluster_frac=0.05, unknown_frac=0.1, phix_frac=0.6)
p = TileProcessor(cfg)
p.encode_configure()
p.upload(t.bcl, t.cif, invert_colormatrix(t.matrix))
p.process()
out = p.encode()
print("encoded", sum(len(v) for v in out.values()), "bytes")
print(len(p.selftest_deflate(bytes(range(256)) * 300)))
p.close()
PY
for tool in memcheck racecheck; do
    timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python /tmp/tsk_san.py > gpurun_out/sanitizer_$tool.log 2>&1
    echo "$tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke ok|encoded" gpurun_out/sanitizer_$tool.log | tail -6
done
