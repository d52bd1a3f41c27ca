#!/bin/bash
run() { echo "== $1"; python tools/profile_step.py 0 3 $2 $3 2>&1 | grep "step 2" | cut -c1-100; }
FS_TUNE_TQ=0 run "room old" furnished_room 16
for f in 16 24 32; do for n in 6 10; do FS_TUNE_TQ_FLUSH=$f FS_TUNE_TQ_NODE_MIN=$n run "room tq flush=$f node_min=$n" furnished_room 16; done; done
for r in 1 2 8; do FS_TUNE_REFILL=$r run "room tq refill=$r" furnished_room 16; done
run "hall tq default" concert_hall 32
FS_TUNE_TQ_FLUSH=32 run "hall tq flush 32" concert_hall 32
