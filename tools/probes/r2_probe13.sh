#!/bin/bash
for sp in 0 2 1; do HVS_SEED_PHASE=$sp python tools/shard_rank_probe.py 8 0 2>&1 | tail -1 | sed "s/^/seed=$sp /"; done
for sp in 0 2; do HVS_SEED_PHASE=$sp HVS_K3_STATS=1 python tools/shard_rank_probe.py 8 0 2>&1 | grep -E "K3 stats|K3 hits" | tail -2 | sed "s/^/seed=$sp /" | cut -c1-300; done
for rep in 1 2; do for sp in 0 2; do
  HVS_SEED_PHASE=$sp timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-configs --no-parity 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('headline seed=$sp', round(l['value']), round(l['ms_per_step'],3), round(l['kernel_ms_per_step']['K3 k_tile_tensor'],3))"
done; done
