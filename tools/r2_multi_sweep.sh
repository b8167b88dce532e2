#!/bin/bash
# shard cost model sweep on N GPUs: $1 = N
N=$1; O=gpurun_out
for qc in 400000 2000000 6000000; do for stp in 16; do
  HVS_SHARD_QCOST=$qc HVS_SHARD_STRIPES=$stp python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 6 --warmup 2 --no-parity > $O/r2s_${N}_${qc}_$stp.json 2> $O/r2s_${N}_${qc}_$stp.err || tail -5 $O/r2s_${N}_${qc}_$stp.err
  python - "$O/r2s_${N}_${qc}_$stp.json" $qc $stp <<'PY'
import json,sys
l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("qcost", sys.argv[2], "stripes", sys.argv[3], "n_gpus", l["n_gpus"], round(l["value"]), round(l["ms_per_step"],3), "per-rank solve ms", [round(r["ms_solve_device"],2) for r in l["per_rank"]], "queries", [r["queries"] for r in l["per_rank"]], "data variant", round(l["data_sharded_variant"]["value"]))
PY
done; done
