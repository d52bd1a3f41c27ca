#!/bin/bash
# A/B timing of env knobs on the furnished room and the hall
run() { echo "== $1"; python tools/profile_step.py 0 3 $2 $3 2>&1 | grep "frequensee\|step 2"; python tools/profile_step.py 1 2 $2 $3 | tail -1; }
export FS_VERBOSE=1
for c in 1 0 3; do
FS_TUNE_COLLAPSE=$c run "room collapse=$c" furnished_room 16
done
for c in 1 0 3; do
FS_TUNE_COLLAPSE=$c run "hall collapse=$c" concert_hall 32
done
