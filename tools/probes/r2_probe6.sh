#!/bin/bash
for W in 8 4; do for ips in 16 8 4 2; do HVS_ITEMS_PER_SM=$ips python tools/shard_rank_probe.py $W 3 2>&1 | tail -1; done; done
