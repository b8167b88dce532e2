#!/bin/bash
# pair-indexed block bitonic sort: GPU tests, then default / headline / medium / selective lines
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for wl in default large medium selective; do timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --no-configs --parity-sample 64 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$wl', round(l['value']), round(l['ms_per_step'],4), {k:round(v,3) for k,v in l['kernel_ms_per_step'].items()}, l['parity']['ok'])"; done
