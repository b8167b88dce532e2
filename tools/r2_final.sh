#!/bin/bash
# round-end confirmation on one GPU: GPU tests, smoke, the driver's bench line, the reference arm
O=gpurun_out
S=$(date +%s)

python bench.py > $O/bench_final.json 2> $O/bench_final.err; echo "bench wall $(( $(date +%s) - S )) s"; S=$(date +%s)
python bench.py --impl reference > $O/bench_final_ref.json 2> $O/bench_final_ref.err; echo "reference wall $(( $(date +%s) - S )) s"
python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[-1])
print('value',round(l['value']),'ms',round(l['ms_per_step'],3),'e2e',round(l['e2e']['value']),'frac',round(l['roofline']['frac'],3),'parity',l['parity']['ok'],l['parity'].get('queries'),'clocks',l['clocks'])
for k,v in l.get('configs',{}).items(): print(' ',k,round(v['value']),round(v['ms_per_step'],3),v.get('roofline',{}).get('frac'),v.get('parity',{}).get('ok'))
r=json.loads(open('gpurun_out/bench_final_ref.json').read().strip().splitlines()[-1]); print('reference',r['value'],r['unit'],r['cpu_baseline'])
PY
