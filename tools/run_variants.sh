for cd in 1 2; do for ch in 2 4 8; do
DLIOM_CHUNKS_DEV=$cd DLIOM_CHUNKS_HOST=$ch python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('dev_chunks=$cd host_chunks=$ch', 'value %.0f ms %.3f | e2e %.0f ms %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step']), {k:v['ms_per_step'] for k,v in d['roofline']['stages'].items()})"
done; done
