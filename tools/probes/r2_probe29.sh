#!/bin/bash
# shard cost model: fixed part 14e6 (default) vs 20e6 / 26e6, every rank's share of the headline batch at W = 8 and 2
for Q in 14000000 20000000 26000000; do
  echo "HVS_SHARD_QCOST=$Q"
  for W in 8 2; do HVS_SHARD_QCOST=$Q python tools/shard_all_ranks.py $W 2>&1 | awk '{print $1,$2,$3,$5,$7,$8}'; done
done
