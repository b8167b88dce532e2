#!/bin/bash
for W in 8 1; do for ctr in 0 65536 131072; do HVS_CT_R=$ctr python tools/shard_rank_probe.py $W 0 2>&1 | tail -1 | sed "s/^/ct_r=$ctr /"; done; done
