#!/bin/bash
export FS_VERBOSE=1
run() { echo "== $1"; python tools/profile_step.py 0 3 $2 $3 2>&1 | grep "rotation\|commit\|step 2" | cut -c1-110; python tools/profile_step.py 1 1 $2 $3 2>&1 | grep "step 0" | cut -c110-230; }
for r in 0 3 8; do
FS_TUNE_ROTATE=$r run "room rotate=$r" furnished_room 16
FS_TUNE_ROTATE=$r run "hall rotate=$r" concert_hall 32
FS_TUNE_ROTATE=$r run "tunnels rotate=$r" mine_tunnels 16
done
