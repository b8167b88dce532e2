#!/bin/bash
for W in 2 4; do for sp in 0 2; do HVS_SEED_PHASE=$sp HVS_TIMELINE=1 python tools/shard_rank_probe.py $W 1 2>&1 | grep -E "world|timeline" | tail -2 | sed "s/^/seed=$sp /" | cut -c1-330; done; done
for wl in type0 type2 type13 medium; do for sp in 0 2; do
  HVS_SEED_PHASE=$sp timeout 300 python bench.py --workload $wl --steps 4 --warmup 2 --no-cpu-baseline --no-configs --no-parity 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$wl seed=$sp', round(l['value']), round(l['ms_per_step'],3), round(l['kernel_ms_per_step']['K3 k_tile_tensor'],3))"
done; done
