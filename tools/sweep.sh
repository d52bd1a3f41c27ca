#!/bin/bash
run() { echo "== $1"; python tools/profile_step.py 0 3 $2 $3 2>&1 | grep "step 2" | cut -c1-100; }
run "room split" furnished_room 16
FS_LIB_PATH=$PWD/audio-pathtracer_b200/lib_var/nosplit/libfrequensee.so run "room nosplit" furnished_room 16
PS_PATHS=1310720 run "hall split" concert_hall 32
PS_PATHS=1310720 FS_LIB_PATH=$PWD/audio-pathtracer_b200/lib_var/nosplit/libfrequensee.so run "hall nosplit" concert_hall 32
