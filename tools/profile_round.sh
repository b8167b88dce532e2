#!/bin/bash
# ncu evidence for profiles/ (run under gpurun, ONE GPU, after the plain bench command has exited 0 without ncu)
O=gpurun_out
A="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity"
$A > $O/prof_plain.log 2>&1 || { echo "plain bench failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r1_bench_auto_launches.csv $A > $O/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_tile_tensor -s 4 -c 1 -f -o $O/r1_k3_bench_full $A > $O/ncu_f.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file $O/r1_traffic_auto.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity > $O/ncu_t1.log 2>&1
E="python bench.py --mode exact --steps 2 --warmup 1 --no-cpu-baseline --no-parity"
ncu --set full --clock-control none --import-source on -k regex:k_tile_ffma -s 4 -c 1 -f -o $O/r1_k2_bench_full $E > $O/ncu_f2.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file $O/r1_traffic_exact.csv python bench.py --mode exact --steps 1 --warmup 1 --no-cpu-baseline --no-parity > $O/ncu_t2.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file $O/r1_traffic_selective.csv python bench.py --workload selective --steps 1 --warmup 1 --no-cpu-baseline --no-parity > $O/ncu_t3.log 2>&1
ls -la $O/*.ncu-rep $O/r1_*.csv
