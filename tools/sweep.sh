#!/bin/bash
run() { echo "== $1"; python tools/profile_step.py 0 3 $2 $3 2>&1 | grep "step 2" | cut -c1-100; python tools/profile_step.py 1 2 $2 $3 2>&1 | grep "step 1" | cut -c110-200; }
FS_TUNE_TQ=0 run "room old" furnished_room 16
for f in 8 16 24 32; do for n in 0 8 16; do FS_TUNE_TQ_FLUSH=$f FS_TUNE_TQ_NODE_MIN=$n run "room tq flush=$f node_min=$n" furnished_room 16; done; done
FS_TUNE_TQ=0 run "hall old" concert_hall 32
run "hall tq default" concert_hall 32
