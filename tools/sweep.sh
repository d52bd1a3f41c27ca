#!/bin/bash
run() { echo "== $1"; python tools/profile_step.py 0 3 $2 $3 2>&1 | grep "step 2" | cut -c1-130; }
run "room any phased" furnished_room 16
FS_TUNE_TQ=2 run "room any queue" furnished_room 16
run "hall any phased" concert_hall 32
FS_TUNE_TQ=2 run "hall any queue" concert_hall 32
run "tunnels any phased" mine_tunnels 16
FS_TUNE_TQ=2 run "tunnels any queue" mine_tunnels 16
