#!/bin/bash
run() { echo "== $1"; python tools/profile_step.py 0 3 $2 $3 2>&1 | grep "step 2" | cut -c1-100; }
run "room base" furnished_room 16
FS_LIB_PATH=$PWD/audio-pathtracer_b200/lib_var/pf/libfrequensee.so run "room prefetch" furnished_room 16
PS_PATHS=1310720 run "hall base" concert_hall 32
PS_PATHS=1310720 FS_LIB_PATH=$PWD/audio-pathtracer_b200/lib_var/pf/libfrequensee.so run "hall prefetch" concert_hall 32
