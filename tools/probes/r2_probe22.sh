#!/bin/bash
# configs[0] (D=10^4, Q=10^2): where do the 0.16 ms go?  launch list + full capture of k_direct
O=gpurun_out
D="python bench.py --workload default --steps 20 --warmup 5 --no-cpu-baseline --no-parity --no-configs"
$D 2>/dev/null | tail -1 > $O/r2p22_default.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r2p22_default_launches.csv python bench.py --workload default --steps 2 --warmup 1 --no-cpu-baseline --no-parity --no-configs > $O/r2p22_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_direct -s 2 -c 1 -f -o $O/r2p22_k4_default python bench.py --workload default --steps 2 --warmup 1 --no-cpu-baseline --no-parity --no-configs > $O/r2p22_f.log 2>&1
ls -la $O/r2p22*
