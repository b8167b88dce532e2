#!/bin/bash
# single-launch cluster sort vs three-launch passes (HVS_SORT_CLUSTER=0): GPU tests, headline + medium, one rank's share of 8
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for C in 1 0; do
export HVS_SORT_CLUSTER=$C
echo "HVS_SORT_CLUSTER=$C"
python tools/shard_rank_probe.py 8 0 2>&1 | tail -1
for wl in large medium; do timeout 300 python bench.py --workload $wl --steps 8 --warmup 3 --no-cpu-baseline --no-configs --parity-sample 64 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$wl', round(l['value']), round(l['ms_per_step'],3), {k:round(v,3) for k,v in l['kernel_ms_per_step'].items()}, l['parity']['ok'], 'index ms', l.get('index_build_ms'))"; done
done
