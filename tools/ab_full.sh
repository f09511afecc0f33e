#!/bin/bash
# A/B of build/lib_base.so vs the current library on ONE box with the full bench line (extras included).
TAG=${1:-abf}
OUT=gpurun_out/${TAG}.log
: > $OUT
LIB=d-liom_b200/libdliom_b200.so
cp $LIB build/lib_new.so
run() {
  python bench.py --steps 60 --cpu-sample 64 2>/dev/null | grep "^{" | python -c "
import json,sys; d=json.loads(sys.stdin.read())
print(round(d['value']), round(d['front_end_only']['value']), 'modeF', round(d['mode_F']['value']), d['mode_F']['stages_ms_per_step'].get('nls_solve'), 'nls_serial', d['roofline']['stages']['nls_solve']['ms_per_step'], 'lat', d['latency']['single_scan_ms'], d['parity_vs_cpu']['all_ok'], d['mode_F']['all_ok'])" >> $OUT
}
for rep in 1 2; do
  cp build/lib_base.so $LIB; echo -n "base " >> $OUT; run
  cp build/lib_new.so $LIB;  echo -n "new  " >> $OUT; run
done
cat $OUT
