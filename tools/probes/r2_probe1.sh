#!/bin/bash
# K4s (warp-per-query) validation: GPU tests, then the selective / default / large workloads with and without it
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r2p1_tests.log 2>&1; tail -3 $O/r2p1_tests.log
for wl in selective default large; do
  for sm in 1 0; do
    HVS_SMALL=$sm python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --no-configs --parity-sample 64 > $O/r2p1_${wl}_small$sm.json 2> $O/r2p1_${wl}_small$sm.err
    python - "$O/r2p1_${wl}_small$sm.json" <<'PY'
import json,sys
l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], round(l["value"]), round(l["ms_per_step"],4), "e2e", round(l["e2e"]["value"]), {k:round(v,4) for k,v in l["kernel_ms_per_step"].items()}, l.get("parity",{}).get("ok"), l["stats"]["launches"])
PY
  done
done
# K3: which lanes take part in a compaction (HVS_K3_PARTMIN), interleaved
for rep in 1 2; do for pm in 100 256 352; do
  HVS_K3_PARTMIN=$pm python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs --no-parity > $O/r2p1_pm${pm}_$rep.json 2> $O/r2p1_pm${pm}_$rep.err
  python - "$O/r2p1_pm${pm}_$rep.json" <<'PY'
import json,sys
l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], round(l["value"]), round(l["ms_per_step"],3), {k:round(v,3) for k,v in l["kernel_ms_per_step"].items()})
PY
done; done
HVS_K3_PARTMIN=256 HVS_K3_STATS=1 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --no-configs > /dev/null 2> $O/r2p1_k3stats_pm256.err
HVS_K3_PARTMIN=256 python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-configs --parity-sample 64 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('pm256 parity', l['parity']['ok'], l['stats']['n_fallback'])"
