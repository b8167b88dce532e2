#!/bin/bash
# strong-scaling runs on N GPUs of one box: $1 = N
N=$1; O=gpurun_out
show() { python - "$1" <<'PY'
import json,sys
l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], "n_gpus", l["n_gpus"], l["scaling"], round(l["value"]), round(l["ms_per_step"],3), "e2e", round(l["e2e"]["value"]), l.get("parity",{}).get("ok"), l.get("parity",{}).get("queries"), "weak", l.get("weak_scaling",{}).get("value"))
for r in l.get("per_rank",[]): print("   ", {k:(round(v,3) if isinstance(v,float) else v) for k,v in r.items()})
PY
}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > $O/r2m_q$N.json 2> $O/r2m_q$N.err || tail -20 $O/r2m_q$N.err
show $O/r2m_q$N.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --variant data > $O/r2m_d$N.json 2> $O/r2m_d$N.err || tail -20 $O/r2m_d$N.err
show $O/r2m_d$N.json
