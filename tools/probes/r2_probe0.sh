#!/bin/bash
# round-2 first probe: round-1 build, uniform vs category-aligned chunks (interleaved), K3 cycle budget, timeline
O=gpurun_out
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline"
for rep in 1 2; do
  $B > $O/r2p0_uniform_$rep.json 2> $O/r2p0_uniform_$rep.err
  HVS_ALIGN_CHUNKS=1 $B > $O/r2p0_align_$rep.json 2> $O/r2p0_align_$rep.err
done
HVS_TIMELINE=1 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity > $O/r2p0_timeline.json 2> $O/r2p0_timeline.err
HVS_ALIGN_CHUNKS=1 HVS_TIMELINE=1 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity > $O/r2p0_timeline_align.json 2> $O/r2p0_timeline_align.err
HVS_K3_STATS=1 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity > $O/r2p0_k3stats.json 2> $O/r2p0_k3stats.err
HVS_ALIGN_CHUNKS=1 HVS_K3_STATS=1 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity > $O/r2p0_k3stats_align.json 2> $O/r2p0_k3stats_align.err
HVS_PLAN_DEBUG=1 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity > /dev/null 2> $O/r2p0_plandbg.err
for f in $O/r2p0_uniform_*.json $O/r2p0_align_*.json; do python - "$f" <<'PY'
import json,sys
l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], round(l["value"]), round(l["ms_per_step"],2), l["kernel_ms_per_step"], l.get("parity",{}).get("ok"))
PY
done
