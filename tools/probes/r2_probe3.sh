#!/bin/bash
# device planner: GPU tests (device planner default), a subset with the host planner, then A/B on the headline workload
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2p3_tests.log 2>&1; tail -3 $O/r2p3_tests.log
HVS_PLAN=host timeout 600 python -m pytest tests -m gpu -x -q -k "fixtures or seeded or ties or sharded" > $O/r2p3_tests_host.log 2>&1; tail -2 $O/r2p3_tests_host.log
show() { python - "$1" <<'PY'
import json,sys
l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], round(l["value"]), round(l["ms_per_step"],4), "e2e", round(l["e2e"]["value"]), {k:round(v,4) for k,v in l["kernel_ms_per_step"].items()}, l.get("parity",{}).get("ok"), l["stats"]["launches"], l["stats"]["n_items_tensor"])
PY
}
for rep in 1 2; do for pl in dev host; do
  HVS_PLAN=$pl timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs --no-parity > $O/r2p3_${pl}_$rep.json 2> $O/r2p3_${pl}_$rep.err; show $O/r2p3_${pl}_$rep.json
done; done
HVS_TIMELINE=1 timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs --no-parity > /dev/null 2> $O/r2p3_timeline.err; tail -2 $O/r2p3_timeline.err
HVS_K3_STATS=1 timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --no-configs > /dev/null 2> $O/r2p3_k3stats.err
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/r2p3_full.json 2> $O/r2p3_full.err; show $O/r2p3_full.json
python - <<'PY'
import json
l=json.loads(open("gpurun_out/r2p3_full.json").read().strip().splitlines()[-1])
print(l.get("parity"))
for k,c in l.get("configs",{}).items(): print(k, round(c["value"]), round(c["ms_per_step"],3), "e2e", round(c["e2e"]["value"]), c["dominant_kernel"], c["roofline"]["frac"] if c["roofline"] else None, c.get("parity",{}).get("ok"))
PY
