#!/bin/bash
# ncu evidence for profiles/ (round 2).  Run under gpurun, ONE GPU, after the plain bench command has exited 0 without ncu.
O=gpurun_out
A="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity --no-configs"
$A > $O/r2prof_plain.log 2>&1 || { echo "plain bench failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_bench_auto_launches.csv $A > $O/r2ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_tile_tensor -s 2 -c 1 -f -o $O/r2_k3_bench_full $A > $O/r2ncu_f.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file $O/r2_traffic_auto.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --no-configs > $O/r2ncu_t1.log 2>&1
E="python bench.py --mode exact --steps 1 --warmup 1 --no-cpu-baseline --no-parity --no-configs"
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file $O/r2_traffic_exact.csv $E > $O/r2ncu_t2.log 2>&1
S="python bench.py --workload selective --steps 2 --warmup 1 --no-cpu-baseline --no-parity --no-configs"
$S > $O/r2prof_plain_sel.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_traffic_selective.csv $S > $O/r2ncu_t3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_small -s 2 -c 1 -f -o $O/r2_k4s_selective_full $S > $O/r2ncu_f3.log 2>&1
ls -la $O/*.ncu-rep $O/r2_*.csv
