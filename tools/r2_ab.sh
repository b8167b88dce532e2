#!/bin/bash
# interleaved A/B of library variants under tools/ab/ on the headline workload: $1 = tag, rest = variant names
tag=$1; shift
O=gpurun_out
lib=$(ls -d project*)/libhvs_b200.so
cp $lib /tmp/orig.so
for rep in 1 2 3; do for v in "$@"; do
  cp tools/ab/$v.so $lib
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs --no-parity > $O/ab_${tag}_${v}_$rep.json 2> $O/ab_${tag}_${v}_$rep.err
  python - "$O/ab_${tag}_${v}_$rep.json" <<'PY'
import json,sys
l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], round(l["value"]), round(l["ms_per_step"],3), {k:round(v,3) for k,v in l["kernel_ms_per_step"].items()})
PY
done; done
cp /tmp/orig.so $lib
