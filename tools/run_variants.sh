# A/B of the front-end pipeline shapes on one box: python bench lines reduced to the numbers that matter.
run() {
  env "$@" python bench.py --steps 20 --warmup 3 --cpu-sample 8 ${BATCH:+--batch $BATCH} 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); e=d['e2e']; print('$*', 'batch', d['config']['scans_per_step_per_gpu'], 'value %.0f ms %.3f | e2e(stream) %.0f ms %.3f | sync %.0f ms %.3f | pcie %s' % (d['value'], d['ms_per_step'], e['value'], e['ms_per_step'], e['sync_call']['value'], e['sync_call']['ms_per_step'], e['pcie_h2d_gbs']), {k:v['ms_per_step'] for k,v in d['roofline']['stages'].items()}, d['parity_vs_cpu']['rmse_m'])"
}
run A=1
run DLIOM_BENCH_CONTEXTS=3
run DLIOM_BENCH_CONTEXTS=4
run DLIOM_BENCH_CONTEXTS=3 DLIOM_CHUNKS_DEV=1
run DLIOM_BENCH_CONTEXTS=4 DLIOM_CHUNKS_DEV=1
BATCH=74 run DLIOM_BENCH_CONTEXTS=4 DLIOM_CHUNKS_DEV=1
