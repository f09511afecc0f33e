#!/bin/bash
# A/B of two builds of the library on ONE box: build/lib_base.so vs the current d-liom_b200/libdliom_b200.so.
# Usage under gpurun: bash tools/ab_frontend.sh <tag> [extra bench args]
TAG=${1:-ab}; shift
OUT=gpurun_out/${TAG}.log
: > $OUT
LIB=d-liom_b200/libdliom_b200.so
cp $LIB build/lib_new.so
run() {
  v=$(python bench.py --steps 60 --no-extras --cpu-sample 128 "$@" 2>/dev/null | grep "^{" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],4), round(d['e2e']['value']), d['parity_vs_cpu']['all_ok'], {k:v['ms_per_step'] for k,v in d['roofline']['stages'].items()})")
  echo "$v" >> $OUT
}
for rep in 1 2; do
  cp build/lib_base.so $LIB; echo -n "base fe   " >> $OUT; run --pairs 0 "$@"
  cp build/lib_new.so $LIB;  echo -n "new  fe   " >> $OUT; run --pairs 0 "$@"
done
cp build/lib_base.so $LIB; echo -n "base full " >> $OUT; run "$@"
cp build/lib_new.so $LIB;  echo -n "new  full " >> $OUT; run "$@"
cat $OUT
