#!/bin/bash
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2p5_tests.log 2>&1; tail -3 $O/r2p5_tests.log
show() { python - "$1" <<'PY'
import json,sys
l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], round(l["value"]), round(l["ms_per_step"],4), "e2e", round(l["e2e"]["value"]), {k:round(v,4) for k,v in l["kernel_ms_per_step"].items()}, l.get("parity",{}).get("ok"), l["stats"]["launches"])
PY
}
for sp in 1 0; do
  HVS_DIRECT_SPLIT=$sp timeout 300 python bench.py --workload default --steps 20 --warmup 5 --no-cpu-baseline --no-configs --parity-sample 100 > $O/r2p5_default_sp$sp.json 2> $O/r2p5_default_sp$sp.err; show $O/r2p5_default_sp$sp.json
done
timeout 300 python bench.py --workload selective --steps 10 --warmup 3 --no-cpu-baseline --no-configs --parity-sample 64 > $O/r2p5_sel.json 2> $O/r2p5_sel.err; show $O/r2p5_sel.json
