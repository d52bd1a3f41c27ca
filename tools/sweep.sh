#!/bin/bash
for m in 1 4 8 12 16; do
  echo -n "refill=$m : "
  FS_TUNE_REFILL=$m timeout 100 python tools/profile_step.py 0 3 | tail -1 | cut -c1-130
done
for m in 12 16; do
  echo -n "node_min=$m : "
  FS_TUNE_NODE_MIN=$m timeout 100 python tools/profile_step.py 0 3 | tail -1 | cut -c1-130
done
