#!/bin/bash
O=gpurun_out
HVS_SEED_PHASE=1 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
show() { python - "$1" <<'PY'
import json,sys
l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], round(l["value"]), round(l["ms_per_step"],3), {k:round(v,3) for k,v in l["kernel_ms_per_step"].items()}, l.get("parity",{}).get("ok"))
PY
}
for rep in 1 2; do for sp in 0 1; do
  HVS_SEED_PHASE=$sp timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs --no-parity > $O/r2p9_seed${sp}_$rep.json 2> $O/r2p9_seed${sp}_$rep.err; show $O/r2p9_seed${sp}_$rep.json
done; done
HVS_SEED_PHASE=1 HVS_TIMELINE=1 timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs --no-parity 2>&1 >/dev/null | grep timeline | tail -1
for sp in 0 1; do HVS_SEED_PHASE=$sp python tools/shard_rank_probe.py 8 0 2>&1 | tail -1 | sed "s/^/seed=$sp /"; done
for sp in 0 1; do HVS_SEED_PHASE=$sp timeout 300 python bench.py --workload type0 --steps 3 --warmup 1 --no-cpu-baseline --no-configs --no-parity > $O/r2p9_t0_seed$sp.json 2>/dev/null; show $O/r2p9_t0_seed$sp.json; done
HVS_SEED_PHASE=1 timeout 300 python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-configs --parity-sample 128 > $O/r2p9_par.json 2>/dev/null; show $O/r2p9_par.json
