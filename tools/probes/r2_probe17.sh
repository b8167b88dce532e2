#!/bin/bash
for W in 8 4; do for ips in 16 32; do HVS_ITEMS_PER_SM=$ips python tools/shard_rank_probe.py $W 0 2>&1 | tail -1; done; done
for ips in 16 32; do HVS_ITEMS_PER_SM=$ips timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-configs --no-parity 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('headline ips=$ips', round(l['value']), round(l['ms_per_step'],3), {k:round(v,3) for k,v in l['kernel_ms_per_step'].items()})"; done
