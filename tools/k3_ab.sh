#!/bin/bash
# developer A/B probe (run under gpurun): $1 = log name, rest = library variants under tools/ab/ (built .so files);
# each variant runs the headline + type-0 workloads twice, interleaved, so that box-to-box differences cancel
log=gpurun_out/$1; shift
lib=$(ls -d project*)/libhvs_b200.so
cp $lib /tmp/orig.so
{ for rep in 1 2; do for v in "$@"; do echo "== $v rep $rep"; cp tools/ab/$v.so $lib; python tools/gpu_dev.py big auto 2>&1 | grep -v "^   ids" | cut -c1-420; done; done; } > $log 2>&1
cp /tmp/orig.so $lib
grep -E "^==|^n=|ok=False" $log | cut -c1-200
