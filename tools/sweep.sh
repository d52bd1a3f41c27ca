#!/bin/bash
run() { echo "== $1"; python tools/profile_step.py 0 3 $2 $3 2>&1 | grep "step 2" | cut -c1-100; }
run "room base" furnished_room 16
FS_LIB_PATH=$PWD/audio-pathtracer_b200/lib_var/mb6/libfrequensee.so run "room 6 CTAs/SM" furnished_room 16
PS_PATHS=1310720 run "hall base" concert_hall 32
PS_PATHS=1310720 FS_LIB_PATH=$PWD/audio-pathtracer_b200/lib_var/mb6/libfrequensee.so run "hall 6 CTAs/SM" concert_hall 32
