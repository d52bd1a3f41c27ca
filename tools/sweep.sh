#!/bin/bash
export PS_NOTIME=1
run() { echo "== $1"; python tools/profile_step.py 0 5 $2 $3 2>&1 | grep "step 4" | cut -c1-60; }
for s in 1 2 3; do FS_TUNE_STREAMS=$s run "room streams=$s" furnished_room 16; done
export PS_PATHS=1310720
for s in 1 2 3; do FS_TUNE_STREAMS=$s run "hall 1.31M streams=$s" concert_hall 32; done
