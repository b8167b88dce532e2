#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for rep in 1 2; do for g in 0 1; do
  HVS_K5_GTHR=$g timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-configs --no-parity 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('headline k5gthr=$g', round(l['value']), round(l['ms_per_step'],3), {k:round(v,3) for k,v in l['kernel_ms_per_step'].items()})"
done; done
for mode in exact; do for g in 0 1; do
  HVS_K5_GTHR=$g timeout 300 python bench.py --mode $mode --steps 2 --warmup 1 --no-cpu-baseline --no-configs --no-parity 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$mode k5gthr=$g', round(l['value']), round(l['ms_per_step'],3), {k:round(v,3) for k,v in l['kernel_ms_per_step'].items()})"
done; done
timeout 600 python bench.py --steps 3 --warmup 1 --no-cpu-baseline --parity-sample 128 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('parity', l['parity']['ok'], l['parity'].get('all_configs_ok'), l['parity']['exact_vs_auto_all_queries'])"
