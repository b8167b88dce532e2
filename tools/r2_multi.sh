#!/bin/bash
# strong-scaling runs on GPUs of one box: arguments = the N values to run (e.g. "8 4 2")
O=gpurun_out
show() { python - "$1" <<'PY'
import json,sys
l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], "n_gpus", l["n_gpus"], l["scaling"], round(l["value"]), round(l["ms_per_step"],3), "e2e", round(l["e2e"]["value"]), "parity", l.get("parity",{}).get("ok"), l.get("parity",{}).get("queries"), "weak", round(l.get("weak_scaling",{}).get("value",0)), "data-sharded", round(l.get("data_sharded_variant",{}).get("value",0)), l.get("data_sharded_variant",{}).get("parity_vs_query_sharded",{}).get("ok"))
for r in l.get("per_rank",[]): print("   ", {k:(round(v,3) if isinstance(v,float) else v) for k,v in r.items()})
PY
}
for N in "$@"; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 > $O/r2m_q$N.json 2> $O/r2m_q$N.err || tail -20 $O/r2m_q$N.err
  show $O/r2m_q$N.json
done
