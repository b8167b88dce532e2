#!/bin/bash
O=gpurun_out
show() { python - "$1" <<'PY'
import json,sys
l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], round(l["value"]), round(l["ms_per_step"],3), {k:round(v,3) for k,v in l["kernel_ms_per_step"].items()}, l.get("parity",{}).get("ok"))
PY
}
for rep in 1 2; do for tg in 384 288 224 176; do
  HVS_K3_TRIG=$tg timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-configs --no-parity > $O/r2p12_tg${tg}_$rep.json 2> /dev/null; show $O/r2p12_tg${tg}_$rep.json
done; done
for tg in 384 224; do HVS_K3_TRIG=$tg timeout 300 python bench.py --workload type13 --steps 4 --warmup 2 --no-cpu-baseline --no-configs --no-parity > $O/r2p12_t13_tg$tg.json 2>/dev/null; show $O/r2p12_t13_tg$tg.json; done
for tg in 384 224; do HVS_K3_TRIG=$tg timeout 300 python bench.py --workload type2 --steps 4 --warmup 2 --no-cpu-baseline --no-configs --no-parity > $O/r2p12_t2_tg$tg.json 2>/dev/null; show $O/r2p12_t2_tg$tg.json; done
HVS_K3_TRIG=224 timeout 300 python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-configs --parity-sample 128 > $O/r2p12_par.json 2>/dev/null; show $O/r2p12_par.json
