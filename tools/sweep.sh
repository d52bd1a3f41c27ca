#!/bin/bash
run() { echo "== $1"; python tools/profile_step.py 0 3 $2 $3 2>&1 | grep "commit\|step 2" | cut -c1-100; python tools/profile_step.py 1 1 $2 $3 2>&1 | grep "step 0" | cut -c110-230; }
run "hall auto" concert_hall 32
FS_TUNE_PLOC_R=3 run "hall r3" concert_hall 32
FS_TUNE_PLOC_R=2 run "hall r2" concert_hall 32
FS_TUNE_PLOC_R=48 run "tunnels r48" mine_tunnels 16
FS_TUNE_PLOC_R=96 run "tunnels r96" mine_tunnels 16
