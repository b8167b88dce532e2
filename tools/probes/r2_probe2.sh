#!/bin/bash
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r2p2_tests.log 2>&1; tail -3 $O/r2p2_tests.log
show() { python - "$1" <<'PY'
import json,sys
l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], round(l["value"]), round(l["ms_per_step"],4), "e2e", round(l["e2e"]["value"]), {k:round(v,4) for k,v in l["kernel_ms_per_step"].items()}, l.get("parity",{}).get("ok"), l["stats"]["launches"])
PY
}
for wl in selective default; do
  python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --no-configs --parity-sample 64 > $O/r2p2_${wl}.json 2> $O/r2p2_${wl}.err; show $O/r2p2_${wl}.json
done
for rep in 1 2; do for ss in 1 0; do
  HVS_K3_SPARSE_SEL=$ss python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs --no-parity > $O/r2p2_ss${ss}_$rep.json 2> $O/r2p2_ss${ss}_$rep.err; show $O/r2p2_ss${ss}_$rep.json
done; done
HVS_K3_STATS=1 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --no-configs > /dev/null 2> $O/r2p2_k3stats.err
python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-configs --parity-sample 128 > $O/r2p2_par.json 2>/dev/null; show $O/r2p2_par.json
python bench.py --workload type0 --steps 3 --warmup 1 --no-cpu-baseline --no-configs --parity-sample 16 > $O/r2p2_type0.json 2>/dev/null; show $O/r2p2_type0.json
