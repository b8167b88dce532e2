#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for W in 8 4 2; do python tools/shard_rank_probe.py $W 0 2>&1 | tail -1; done
for wl in medium large type13; do timeout 300 python bench.py --workload $wl --steps 6 --warmup 3 --no-cpu-baseline --no-configs --parity-sample 64 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$wl', round(l['value']), round(l['ms_per_step'],3), {k:round(v,3) for k,v in l['kernel_ms_per_step'].items()}, l['parity']['ok'], int(l['stats']['n_items_tensor']))"; done
