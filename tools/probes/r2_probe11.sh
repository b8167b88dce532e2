#!/bin/bash
O=gpurun_out
show() { python - "$1" <<'PY'
import json,sys
l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], round(l["value"]), round(l["ms_per_step"],3), {k:round(v,3) for k,v in l["kernel_ms_per_step"].items()}, l.get("parity",{}).get("ok"))
PY
}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for rep in 1 2 3; do for lh in 0 1; do
  HVS_K3_L2HINTS=$lh timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-configs --no-parity > $O/r2p11_lh${lh}_$rep.json 2> /dev/null; show $O/r2p11_lh${lh}_$rep.json
done; done
for lh in 0 1; do HVS_K3_L2HINTS=$lh timeout 300 python bench.py --workload type0 --steps 4 --warmup 2 --no-cpu-baseline --no-configs --no-parity > $O/r2p11_t0_lh$lh.json 2>/dev/null; show $O/r2p11_t0_lh$lh.json; done
HVS_K3_L2HINTS=1 timeout 300 python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-configs --parity-sample 128 > $O/r2p11_par.json 2>/dev/null; show $O/r2p11_par.json
# f4 question: would a first pass at twice the tensor rate (FP8 operands) help?  4 of 7 k-steps issued (results wrong), instrumented build both times
for hm in 0 1; do HVS_K3_STATS=1 HVS_K3_HALF_MMA=$hm timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs --no-parity 2> $O/r2p11_hm$hm.err | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('half_mma=$hm (instrumented build) K3 ms', round(l['kernel_ms_per_step']['K3 k_tile_tensor'],3))"; grep "K3 stats" $O/r2p11_hm$hm.err | tail -1 | cut -c1-330; done
for hm in 0 1; do HVS_K3_STATS=1 HVS_K3_HALF_MMA=$hm timeout 300 python bench.py --workload type0 --steps 2 --warmup 1 --no-cpu-baseline --no-configs --no-parity 2> $O/r2p11_t0_hm$hm.err | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('type0 half_mma=$hm (instrumented build) K3 ms', round(l['kernel_ms_per_step']['K3 k_tile_tensor'],3))"; grep "K3 stats" $O/r2p11_t0_hm$hm.err | tail -1 | cut -c1-330; done
