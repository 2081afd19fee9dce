cd tailseeker_b200
for rep in 1 2; do
for v in prev new; do
  cp lib_$v.bin libtailseeker_b200.so
  (cd .. && python bench.py --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "
import sys, json
d=json.loads(sys.stdin.read()); s=d['roofline']['stages_ms_per_tile']; print('$v', round(d['value']/1e6,2), round(d['ms_per_step'],2), round(s['normalise_score_pack'],3), round(s['fair_sampling_fixups'],3))")
done; done
