#!/bin/bash
# Env-knob A/B of the front-end step (device-resident, no exchange) on ONE box. Usage under gpurun:
#   bash tools/sweep_frontend.sh <tag> "name ENV=val ..." "name2 ENV=val" ...
TAG=${1:-sweep}; shift
OUT=gpurun_out/${TAG}.log
: > $OUT
run() {
  name=$1; shift
  v=$(env "$@" python bench.py --steps 60 --no-extras --pairs 0 --cpu-sample 128 2>/dev/null | grep "^{" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],4), d['parity_vs_cpu']['all_ok'], {k:v['ms_per_step'] for k,v in d['roofline']['stages'].items()})")
  echo "$name $v" >> $OUT
}
for spec in "$@"; do run $spec; done
cat $OUT
